# Round-2 GPU session script (run under gpurun from the repo root):  bash tests/_r02_run.sh <stage> [...]
# stages: tests | smoke | bench | traffic | launches | full
set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
TAG=${TAG:-r02}
for stage in "$@"; do
case $stage in
tests)
  timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
  grep "^\[parity\]" gpurun_out/${TAG}_pytest.log > gpurun_out/${TAG}_parity_lines.txt ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -4 gpurun_out/${TAG}_smoke.log ;;
bench)
  : > gpurun_out/${TAG}_bench_lines.jsonl
  timeout 900 python bench.py >> gpurun_out/${TAG}_bench_lines.jsonl 2> gpurun_out/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 >> gpurun_out/${TAG}_bench_lines.jsonl 2>> gpurun_out/${TAG}_bench.err
  for w in ${WORKLOADS:-c1 c3 c4 c5w c5 c2v}; do timeout 900 python bench.py --workload $w --steps 50 --warmup 5 >> gpurun_out/${TAG}_bench_lines.jsonl 2>> gpurun_out/${TAG}_bench.err; done
  cut -c1-400 gpurun_out/${TAG}_bench_lines.jsonl; tail -5 gpurun_out/${TAG}_bench.err ;;
traffic)
  for w in ${WORKLOADS:-c2 c1 c3 c4 c5w}; do
    timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/${TAG}_traffic_$w.csv python bench.py --workload $w --ncu 1 --clouds uniform > gpurun_out/${TAG}_traffic_$w.log 2>&1
    tail -1 gpurun_out/${TAG}_traffic_$w.log
  done ;;
full)
  timeout 900 ncu --profile-from-start off -k regex:tc_fused_kernel --set full --clock-control none --import-source on \
    -o gpurun_out/${TAG}_c2_fused -f python bench.py --workload c2 --ncu 1 --clouds uniform > gpurun_out/${TAG}_full_c2.log 2>&1
  tail -2 gpurun_out/${TAG}_full_c2.log; ls -la gpurun_out/*.ncu-rep ;;
esac
done
