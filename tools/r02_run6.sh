set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -15 gpurun_out/r02j_pytest.log
timeout 300 python tools/kbench.py --grids g1,g2,g3,g5 --algos 0,6 --iters 20 > gpurun_out/r02j_kbench_tma.txt 2>&1
timeout 300 python tools/kbench.py --grids g2,g5 --algos 0,6,2 --iters 10 --size 128 --batch 8 >> gpurun_out/r02j_kbench_tma.txt 2>&1
cat gpurun_out/r02j_kbench_tma.txt

