set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01zz_pytest.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01zz_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r01zz_bench_reference.json 2> gpurun_out/r01zz_bench_reference.err; echo "ref rc=$?"
python bench.py --steps 30 --warmup 5 > gpurun_out/r01zz_bench.json 2> gpurun_out/r01zz_bench.err; echo "bench rc=$?"
timeout 700 python tools/bench_configs.py --frames 1000 --iters 4 > gpurun_out/r01zz_configs_2_3_4.jsonl 2> gpurun_out/r01zz_configs.err; echo "cfg rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01zz_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r01zz_ncu_launches.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_warp_fused|k_raster' -c 4 -o gpurun_out/prof_r01zz_gt -f python tools/prof_once.py gt > gpurun_out/prof_r01zz_gt.log 2>&1; echo "ncu gt rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_screen|k_peak|k_limbs|k_assemble|k_axis' -c 14 -o gpurun_out/prof_r01zz_dec -f python tools/prof_once.py decode > gpurun_out/prof_r01zz_dec.log 2>&1; echo "ncu dec rc=$?"
tail -2 gpurun_out/r01zz_pytest.log
