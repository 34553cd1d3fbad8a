set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02a_topo.txt; numactl -H >> gpurun_out/r02a_topo.txt 2>&1; free -g >> gpurun_out/r02a_topo.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; echo "ref rc=$?"
for v in "" variants/libvolstn_b200_ctas5.so variants/libvolstn_b200_ctas4.so; do
  echo "== lib ${v:-default (ctas6)}" >> gpurun_out/r02a_kbench_ctas.txt
  if [ -n "$v" ]; then export VOLSTN_B200_LIB=$GRAFT_REPO_ROOT/alignet_b200/csrc/$v; else unset VOLSTN_B200_LIB; fi
  timeout 300 python tools/kbench.py --grids g1,g2,g3 --algos 0 --iters 20 >> gpurun_out/r02a_kbench_ctas.txt 2>&1
done
unset VOLSTN_B200_LIB
for g in g2 g3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sampler -o gpurun_out/r02a_ncu_$g -f python tools/prof_once.py --what packed --grid $g > gpurun_out/r02a_ncu_$g.log 2>&1
done
ls -la gpurun_out | tail -20
