mkdir -p gpurun_out
rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_decode.txt
cat gpurun_out/t_decode.txt
timeout 300 python tools/decode_ab.py 256 >> gpurun_out/decode_ab.jsonl 2>gpurun_out/ab_err.txt
timeout 600 python tools/fuzz_parity.py 0 200 81 > gpurun_out/fuzz_s.json 2> gpurun_out/fuzz_s_err.txt
cat gpurun_out/fuzz_s.json
