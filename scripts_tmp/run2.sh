set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
timeout 600 python -m pytest tests/test_multirank.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s3_mr_pytest.log
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/s3_b2_f64.json 2> gpurun_out/s3_b2_f64.err
MEEPCU_NO_OVERLAP=1 timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/s3_b2_f64_noov.json 2> gpurun_out/s3_b2_f64_noov.err
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 5 --dtype f32 > gpurun_out/s3_b2_f32.json 2> gpurun_out/s3_b2_f32.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s3_b1_f64.json 2> gpurun_out/s3_b1_f64.err
MEEPCU_NO_PIN=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s3_b1_f64_nopin.json 2> gpurun_out/s3_b1_f64_nopin.err
free -g > gpurun_out/s3_free.txt
