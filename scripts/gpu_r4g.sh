#!/bin/bash
# default bench (headline + config-3 / config-4 extras with their collectives) at N GPUs, launched as the driver launches it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r4g_bench_n$N.json 2> gpurun_out/r4g_bench_n$N.err; echo "bench N=$N exit=$? in $(( $(date +%s) - S )) s"
grep '^{' gpurun_out/r4g_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
for k,v in d.get('extra',{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk not in('workload','metric','dtype')})
"; tail -3 gpurun_out/r4g_bench_n$N.err
