#!/bin/bash
# Multi-GPU records (GPU box with N GPUs): weak + strong scaling of configs[1], the complex-sharded set (configs[3]) and the
# ligand-sharded screen (configs[4]).  usage: bash scripts/scale_r2.sh N [complexes] [ligands]
set -u
cd "$(dirname "$0")/.."
N=${1:-2}; C=${2:-$((8 * N))}; L=${3:-$((32 * N))}
OUT=gpurun_out/scale_r2; mkdir -p $OUT
run() { # name, args...
  local name=$1; shift
  if [ "$N" = "1" ]; then timeout -s KILL 900 python bench.py --gpus 1 "$@" > $OUT/${name}_n$N.json 2> $OUT/${name}_n$N.err
  else timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > $OUT/${name}_n$N.json 2> $OUT/${name}_n$N.err; fi
  echo "== $name N=$N rc=$?"; tail -n 1 $OUT/${name}_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','ms_per_step','scaling','n_gpus')}, 'e2e', d.get('e2e',{}).get('value'), d.get('clocks',{}).get('sm_mhz'))" 2>/dev/null || tail -n 3 $OUT/${name}_n$N.err
}
run weak --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-grade
run strong --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-grade --scaling strong
run pdbbind --workload pdbbind_synth --complexes $C
run screen --workload screen --complexes $L
