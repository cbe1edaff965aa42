cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sh in 0/1 3/8; do echo "== shard $sh"; FB_SHARD=$sh timeout 300 python tools/variant_bench.py run 2>&1 | tail -2; done
