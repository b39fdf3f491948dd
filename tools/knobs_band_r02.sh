#!/bin/bash
# shadow-sweep chunk schemes on ONE 1/8 band share of the C4 frame (what each GPU of an 8-GPU run renders)
export PROBE_BANDS=8,0,8
for env in "" "TRACER_GEO=64,1,2,3,4,6,8,12,16,24,32,48,64" "TRACER_GEO=64,1,2,3,4,5,6,8,10,12,16,20,24,32,40,48,64" "TRACER_GEO=64,1,2,4,8,16,32,48,64" "TRACER_GEO=128,1,2,3,4,6,8,12,16,24,32,48,64,96,128" "TRACER_CHUNKS=32" "TRACER_CHUNKS=16" "TRACER_ITEMS_PER_CTA=8" "TRACER_ITEMS_PER_CTA=48" "TRACER_MIN_TILES=2"; do
  echo "== $env"
  env $env timeout 120 python tools/probe.py nopeak 1000000,3840,2160,4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_total','ms_primary','ms_shadow','tests_shadow','tests_shadow_ref')})"
done
