mkdir -p gpurun_out
echo "== nb256 ring2"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=67108864 timeout 600 python scripts/trace_fused.py 2>&1 | tail -12
echo "== nb128 ring2"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=33554432 timeout 600 python scripts/trace_fused.py 2>&1 | tail -12
echo "== nb128 ring4"; NNCONV_RING=4 NNCONV_B200_Y_BYTES=67108864 timeout 600 python scripts/trace_fused.py 2>&1 | tail -12
