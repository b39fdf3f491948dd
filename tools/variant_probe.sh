#!/bin/bash
# development: time a variant build of the library (tools/variants/*.so) against the in-tree one, full C4 frame and one 1/8 band share
run() { python tools/probe.py nopeak 1000000,3840,2160,4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_total','ms_primary','ms_shadow','tests_shadow')})"; }
for v in "" $@; do
  if [ -n "$v" ]; then cp tools/variants/$v esctp1raytracer_b200/libtracer_cuda.so; fi
  echo "== ${v:-in-tree} full"; run
  echo "== ${v:-in-tree} band 1/8"; PROBE_BANDS=8,0,8 run
done
