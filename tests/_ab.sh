# same-box A/B of two builds of libp3tok.so:  bash tests/_ab.sh "<workloads>" [extra bench args]
# A = p3tok/libp3tok_prev.so (previous commit), B = p3tok/libp3tok.so (working tree)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
PKG=adapting-2d-vits-for-3d-point-cloud-understanding_b200/p3tok
for w in $1; do
  for lib in libp3tok_prev.so libp3tok.so libp3tok_prev.so libp3tok.so; do
    P3TOK_LIB=$PWD/$PKG/$lib python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-extra --clouds uniform ${@:2} 2>>gpurun_out/ab.err | \
      python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$w', '$lib', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'embed_ms', d['stage_ms_per_step'].get('embed'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
  done
done
