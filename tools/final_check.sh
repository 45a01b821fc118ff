timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f1_tests.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/f1_ref.json 2> gpurun_out/f1_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/f1_tests.log
