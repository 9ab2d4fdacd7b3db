M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -c 1500 --csv --log-file gpurun_out/launches_forceshard.csv python bench.py --force-shard --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-fp32 > gpurun_out/ncu_fs.log 2>&1
tail -c 200 gpurun_out/ncu_fs.log
