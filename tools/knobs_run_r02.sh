#!/bin/bash
# development: run-length cap of the shadow sweeps' self-scheduling (pairs per run), full C4 frame and a 1/8 band share
export TRACER_SHADOW_DIAG=1
for env in "TRACER_RUN_PAIRS=3000000" "TRACER_RUN_PAIRS=1000000000000" "TRACER_RUN_PAIRS=1500000" "TRACER_RUN_PAIRS=6000000" "TRACER_RUN_PAIRS=3000000 TRACER_ITEMS_PER_CTA=96"; do
  for b in "" "8,0,8"; do
    echo "== $env bands [$b]"
    env $env PROBE_BANDS=$b timeout 120 python tools/probe.py nopeak 1000000,3840,2160,4 2> gpurun_out/diag.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_total','ms_primary','ms_shadow')})"; tail -1 gpurun_out/diag.err
  done
done
