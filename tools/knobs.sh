#!/bin/bash
# tools/knobs.sh: sensitivity of the configs[1] graph step to the stream-level scheduling knobs (one short bench run each)
run() { tag=$1; shift; env "$@" timeout 100 python bench.py --no-cpu --no-clocks --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],3), 'ms/step  tracer', round(d['roofline']['kernel_ms_per_step'],3))"; }
run baseline A=1
run main_priority_high IRONB_GRAPH_MAIN_PRIORITY=-1
run no_wgrad_stream IRONB_WGRAD_STREAM=0
run no_eik_stream IRONB_GRAPH_EIK_STREAM=0
run no_mat_streams IRONB_GRAPH_MAT_STREAMS=0
