set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do timeout 60 tools/probe/tma_probe $v >> gpurun_out/r02c_tma_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/r02c_tma_probe.txt; done
cat gpurun_out/r02c_tma_probe.txt
timeout 300 python tools/e2e_diag.py > gpurun_out/r02c_e2e_diag.txt 2>&1
cat gpurun_out/r02c_e2e_diag.txt
timeout 600 python -m pytest tests/test_gpu_host_api.py tests/test_gpu_lua_glue.py -q -x > gpurun_out/r02c_pytest_host.log 2>&1; tail -5 gpurun_out/r02c_pytest_host.log
