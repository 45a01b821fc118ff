timeout 300 python -m pytest tests/test_generation_gpu.py -x -q -k "lanes" > gpurun_out/t4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t4_tests.log
for n in 1 2 4; do VACNIC_DECODE_LANES=$n timeout 200 python bench.py --workload infer --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/lanes_$n.json 2> gpurun_out/lanes_$n.err; echo "lanes $n rc=$?"; done
tail -4 gpurun_out/t4_tests.log
