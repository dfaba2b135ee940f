# N-GPU check of bench.py exactly as the driver launches it (torchrun, NCCL): the cfg3 line (with its cfg4 section), cfg4 as the main
# workload, the sharded evaluate_pose against the single-process table with the real model, and the reference arm under torchrun.
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 scripts/gpu_eval_nccl.py > gpurun_out/eval_nccl_n$N.log 2>&1; echo "eval nccl n=$N exit $?"; grep -E "world|NCCL EVAL" gpurun_out/eval_nccl_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N exit $?"; cat gpurun_out/bench_n$N.json | cut -c1-400; tail -5 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --workload cfg4 > gpurun_out/bench_cfg4_n$N.json 2> gpurun_out/bench_cfg4_n$N.err; echo "bench cfg4 n=$N exit $?"; cat gpurun_out/bench_cfg4_n$N.json | cut -c1-600; tail -5 gpurun_out/bench_cfg4_n$N.err
if [ "${2:-}" = "ref" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref n=$N exit $?"; cat gpurun_out/bench_ref_n$N.json | cut -c1-300
fi
# data-parallel training steps (HeadTrainer + NCCL all-reduce of the flat gradient buffer) == single-process steps on the whole batch
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/gpu_train_ddp.py > gpurun_out/train_ddp_n$N.log 2>&1; echo "train ddp n=$N exit $?"; grep "\[ddp\]" gpurun_out/train_ddp_n$N.log
