# A/B runs at N = 2: what slows a rank down when its neighbour is busy?  usage: bash tools/gpu_n2test.sh
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p gpurun_out/multi
for v in gloo lazy; do
  case $v in
    gloo) E="HRB_DIST_BACKEND=gloo";; lazy) E="HRB_DIST_LAZY=1";;
  esac
  env $E $TR --master-port 29539 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/multi/t_$v.json 2> gpurun_out/multi/t_$v.err
  echo "$v: $(grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/multi/t_$v.json | head -1)"; tail -n 1 gpurun_out/multi/t_$v.err | cut -c1-200
done
