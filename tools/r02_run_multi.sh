set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02k_topo8.txt 2>&1; lscpu | grep -i "^CPU(s)\|NUMA\|Model name" >> gpurun_out/r02k_topo8.txt; free -g >> gpurun_out/r02k_topo8.txt
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02k_bench_n$n.json 2> gpurun_out/r02k_bench_n$n.err; echo "bench n=$n rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 5 --warmup 2 > gpurun_out/r02k_bench_ref_n8.json 2> gpurun_out/r02k_bench_ref_n8.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02k_pytest_multi.log 2>&1; tail -3 gpurun_out/r02k_pytest_multi.log
ls -la gpurun_out | grep r02k
