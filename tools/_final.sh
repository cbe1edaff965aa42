# round-2 final: GPU test suite, smoke, then the measurement campaign of the final build
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_gputest_full.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2_smoke.txt
bash tools/_campaign.sh
