#!/bin/bash
# stand-alone HBM rooflines at V = 2M: event-timed run, then the same launches under ncu for the DRAM bytes
mkdir -p gpurun_out
timeout 300 python scripts/roofline_standalone.py --iters 10 --json gpurun_out/rooflines_standalone.json 2>&1 | tail -12
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:'word_gather|rows_scatter|rows_gather|dgrad' --csv --log-file gpurun_out/rooflines_ncu.csv \
   python scripts/roofline_standalone.py --iters 2 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/rooflines_ncu.csv | head -12
