# Round-end evidence run (one GPU): tests, smoke, both bench arms, the ncu launch list of the bench command and the
# ncu --set full captures of the GT and decode kernels.  usage: bash tools/round_end_capture.sh r02z
tag=${1:-r02z}
set -x
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_warp_fused|k_raster' -c 4 -o gpurun_out/prof_${tag}_gt -f python tools/prof_once.py gt > gpurun_out/prof_${tag}_gt.log 2>&1; echo "ncu gt rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_raster' -c 2 -o gpurun_out/prof_${tag}_crowd -f python tools/prof_once.py gt 20 64 > gpurun_out/prof_${tag}_crowd.log 2>&1; echo "ncu crowd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_screen|k_peak|k_limbs|k_assemble|k_axis' -c 14 -o gpurun_out/prof_${tag}_dec -f python tools/prof_once.py decode > gpurun_out/prof_${tag}_dec.log 2>&1; echo "ncu dec rc=$?"
tail -2 gpurun_out/${tag}_pytest.log
