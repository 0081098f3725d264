run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/r2_scale_$name.json 2> gpurun_out/r2_scale_$name.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" > gpurun_out/r2_scale_$name.json 2> gpurun_out/r2_scale_$name.err; fi
  echo "$name rc=$?"
}
run c2_n2 2 --steps 20 --warmup 3
run c2_n4 4 --steps 20 --warmup 3
run c2_n8 8 --steps 20 --warmup 3
run c2_n8_nccl 8 --steps 20 --warmup 3 --allreduce nccl
run c4_n1 1 --config c4 --steps 10 --warmup 3
run c4_n8 8 --config c4 --steps 10 --warmup 3
run c5_n1 1 --config c5 --steps 10 --warmup 3
run c5_n8 8 --config c5 --steps 10 --warmup 3
run c5_n8_nccl 8 --config c5 --steps 10 --warmup 3 --allreduce nccl
