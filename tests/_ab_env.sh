# same-box A/B of an environment switch:  bash tests/_ab_env.sh "<workloads>" VAR=a VAR=b [extra bench args]
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for w in $1; do
  for rep in 1 2; do for kv in "$2" "$3"; do
    env $kv python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-extra --clouds both ${@:4} 2>>gpurun_out/ab.err | \
      python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$w', '$kv', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'clustered %.0f' % d['other_clouds']['value'], 'stages', d['stage_ms_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
  done; done
done
