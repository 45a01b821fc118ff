TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"
timeout 200 $TR --master-port 29512 tools/dp_p2p_check.py --steps 2 --modes nccl,p2p > gpurun_out/n2_p2p_check.log 2>&1; echo "dp_p2p_check rc=$?"
timeout 200 $TR --master-port 29513 tools/dp_check.py > gpurun_out/n2_dp_check.log 2>&1; echo "dp_check eager rc=$?"
DP_CHECK_GRAPH_ONLY=1 timeout 200 $TR --master-port 29514 tools/dp_check.py >> gpurun_out/n2_dp_check.log 2>&1; echo "dp_check graph rc=$?"
grep "rank" gpurun_out/n2_dp_check.log | tail -4; tail -5 gpurun_out/n2_p2p_check.log
