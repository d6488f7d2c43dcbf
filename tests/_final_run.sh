set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -4 gpurun_out/final_smoke.log
: > gpurun_out/r01b_bench_lines.jsonl
python bench.py >> gpurun_out/r01b_bench_lines.jsonl 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 5 --warmup 1 >> gpurun_out/r01b_bench_lines.jsonl 2>> gpurun_out/final_bench.err
for w in c1 c3 c4 c5 c2v; do python bench.py --workload $w --steps 50 --warmup 5 >> gpurun_out/r01b_bench_lines.jsonl 2>> gpurun_out/final_bench.err; done
cut -c1-220 gpurun_out/r01b_bench_lines.jsonl
K="regex:attention|ln_rows_kernel|norm_max_kernel|tc_linear_kernel"
ncu -k "$K" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_vit4.csv python tests/_vit_time.py > gpurun_out/ncu_vit4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --eager > gpurun_out/ncu_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --eager > gpurun_out/ncu_c2b.log 2>&1
ncu -k regex:tc_linear_kernel --launch-skip 14 --launch-count 1 --set full --clock-control none --import-source on -o gpurun_out/prof_r01_vit_fc1d -f python tests/_vit_time.py > gpurun_out/ncu_full_vit1.log 2>&1
ncu -k regex:attention --launch-skip 3 --launch-count 1 --set full --clock-control none --import-source on -o gpurun_out/prof_r01_vit_attn -f python tests/_vit_time.py > gpurun_out/ncu_full_vit2.log 2>&1
ls -la gpurun_out/*.ncu-rep
