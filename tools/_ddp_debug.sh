timeout 200 python -m pytest tests/test_gpu_sampling.py -x -q -s -k data_parallel 2>&1 | tail -40
for ov in 1 0; do
  SPVD_DDP_OVERLAP=$ov python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_bench_ddp.py 6 2>&1 | tail -2
done
python tools/train_bench_ddp.py 6 2>&1 | tail -1
