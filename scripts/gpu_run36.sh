#!/bin/bash
cd /root/repo
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in j['kernel_ms_per_step'].items() if v['ms']>0}, j['clocks']['sm_mhz'])
"
}
for rep in 1 2; do
run default X=1
run y96MB NNCONV_B200_Y_BYTES=100663296
run y24MB NNCONV_B200_Y_BYTES=25165824
run ring4x128 NNCONV_RING=4 NNCONV_B200_Y_BYTES=67108864
run ybn128 NNCONV_Y_BLOCKN=128
run stages6 NNCONV_APPLY_STAGES=6
done
