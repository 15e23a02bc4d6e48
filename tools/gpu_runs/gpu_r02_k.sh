set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/k_pytest.txt
tail -3 gpurun_out/k_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/k_smoke.txt 2>&1
tail -2 gpurun_out/k_smoke.txt
