TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p gpurun_out/multi
for v in nobind clk200 n1tr; do
  case $v in
    base) E="";; nobind) E="HRB_NO_BIND=1";; clk200) E="HRB_CLOCK_MS=200";; n1tr) E="";;
  esac
  if [ $v = n1tr ]; then
    env $E python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/multi/t_$v.json 2> gpurun_out/multi/t_$v.err
  else
    env $E $TR --master-port 29539 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/multi/t_$v.json 2> gpurun_out/multi/t_$v.err
  fi
  echo "$v: $(grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/multi/t_$v.json | head -1)"; tail -n 1 gpurun_out/multi/t_$v.err | cut -c1-200
done

