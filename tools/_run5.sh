mkdir -p gpurun_out
rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_decode.txt
cat gpurun_out/t_decode.txt
RMPE_SCREEN_CULL=1 timeout 300 python tools/decode_ab.py 256 >> gpurun_out/decode_ab.jsonl 2>gpurun_out/ab_err_1.txt
RMPE_SCREEN_CULL=1 RMPE_SS_GROUP=6 RMPE_MS_GROUP=9 timeout 300 python tools/decode_ab.py 256 >> gpurun_out/decode_ab.jsonl
RMPE_SCREEN_CULL=1 RMPE_SS_GROUP=4 RMPE_MS_GROUP=12 timeout 300 python tools/decode_ab.py 256 >> gpurun_out/decode_ab.jsonl
RMPE_SCREEN_CULL=1 RMPE_SS_GROUP=9 RMPE_MS_GROUP=18 timeout 300 python tools/decode_ab.py 256 >> gpurun_out/decode_ab.jsonl
timeout 600 python tools/fuzz_parity.py 0 200 6 > gpurun_out/fuzz_cull.json 2> gpurun_out/fuzz_cull_err.txt
cat gpurun_out/fuzz_cull.json; tail -3 gpurun_out/fuzz_cull_err.txt
