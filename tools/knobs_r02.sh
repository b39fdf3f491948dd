#!/bin/bash
# shadow-sweep knob sweep at C4 (development): chunk count, items per CTA, smallest slice
for env in "" "TRACER_CHUNKS=32" "TRACER_CHUNKS=48" "TRACER_ITEMS_PER_CTA=12" "TRACER_ITEMS_PER_CTA=48" "TRACER_MIN_TILES=2" "TRACER_MIN_TILES=4" "TRACER_RAYS=16"; do
  echo "== $env"
  env $env timeout 120 python tools/probe.py nopeak 1000000,3840,2160,4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_total','ms_primary','ms_shadow','tests_shadow','strict_evals')})"
done
