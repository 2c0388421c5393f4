#!/bin/bash
O=gpurun_out; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum
python tools/dram_write_probe.py > $O/r02j_plain.log 2>&1 &&
ncu --metrics $M --clock-control none -s 2 -c 4 --csv --log-file $O/r02j_dram_write_probe.csv python tools/dram_write_probe.py > /dev/null 2>&1; echo "ncu exit $?"
cat $O/r02j_dram_write_probe.csv | tail -24
