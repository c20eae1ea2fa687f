mkdir -p gpurun_out
NNCONV_B200_Y_BYTES=100663296 timeout 600 python scripts/trace_apply.py 2>&1 | tail -60
