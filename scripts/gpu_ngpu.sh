# N-GPU bench line (N = $NG) through torchrun, as the driver launches it
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NG:-2}; T=${TAG:-ngpu}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
tail -c 800 gpurun_out/${T}_bench_${N}gpu.err; python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_${N}gpu.json"))
print({k:d.get(k) for k in ("value","ms_per_step","value_with_gather","value_compute_only")})
print("strong", d.get("strong")); print("e2e", d.get("e2e")); print("sharded", json.dumps(d.get("sharded"))[:1500]); print(d["config"].get("gather_check"), d["config"].get("timed_call"))
PY
