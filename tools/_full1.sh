set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_gputest_full.txt
python tools/accuracy_report.py > gpurun_out/r2_accuracy_d.txt 2>&1; cut -c1-250 gpurun_out/r2_accuracy_d.txt
