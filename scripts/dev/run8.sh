set +e
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02e_scale.jsonl
for N in 2 4 8; do $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 >> gpurun_out/r02e_scale.jsonl; done
python scripts/multi_one_process.py --gpus 8 --iters 5 2>/dev/null | tail -1 > gpurun_out/r02e_multi_one_process.jsonl
python scripts/multi_one_process.py --gpus 4 --iters 5 2>/dev/null | tail -1 >> gpurun_out/r02e_multi_one_process.jsonl
$TR --nproc-per-node 8 --master-port 29620 scripts/secondary_multi.py --iters 5 2>/dev/null | tail -1 > gpurun_out/r02e_secondary.jsonl
$TR --nproc-per-node 8 --master-port 29621 scripts/decode_multi.py --iters 5 2>/dev/null | tail -1 > gpurun_out/r02e_decode_multi.jsonl
$TR --nproc-per-node 4 --master-port 29622 scripts/decode_multi.py --iters 5 2>/dev/null | tail -1 >> gpurun_out/r02e_decode_multi.jsonl
$TR --nproc-per-node 8 --master-port 29623 scripts/decode_multi.py --iters 5 --quality 100 2>/dev/null | tail -1 >> gpurun_out/r02e_decode_multi.jsonl
$TR --nproc-per-node 8 --master-port 29624 scripts/batch_1080p.py --images 16384 2>/dev/null | tail -1 > gpurun_out/r02e_batch.jsonl
./scripts/facade_bench.sh 8320 40000 8 > gpurun_out/r02e_facade.jsonl 2>/dev/null
./scripts/facade_bench.sh 8320 40000 1 >> gpurun_out/r02e_facade.jsonl 2>/dev/null
python -m pytest tests/test_gpu_strips.py -q 2>&1 | tail -1
for f in gpurun_out/r02e_*.jsonl; do echo == $f; cut -c1-420 $f; done
