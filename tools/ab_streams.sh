timeout 400 python -m pytest tests/test_trainer_gpu.py -x -q -k "side_streams" > gpurun_out/t2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t2_tests.log
for c in "0 0" "1 0" "0 1" "1 1"; do set -- $c; VACNIC_SIDE_STREAM=$1 VACNIC_GUIDE_STREAM=$2 timeout 200 python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --no-gpu-eager --no-infer > gpurun_out/ab_$1_$2.json 2> gpurun_out/ab_$1_$2.err; echo "ab $c rc=$?"; done
tail -4 gpurun_out/t2_tests.log
