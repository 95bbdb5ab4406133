run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 30 --warmup 3 --no-cpu-baseline --no-ndt --no-configs0 --no-sharded --no-batch 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['value'], d['ms_per_step'], [round(x,3) for x in d['ms_per_step_per_rank']], d['clocks'], d['e2e']['value'], d['e2e'].get('host_numa'))"; }
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
run 29601 default
RSPCL_SYNC=spin run 29602 spin
BENCH_PIN=1 RSPCL_SYNC=spin run 29603 spin+pin
