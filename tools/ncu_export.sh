#!/bin/bash
# usage: tools/ncu_export.sh <name> <kernel regex> <skip> <count> <cmd...>
# Captures `ncu --set full` for the matching launches, exports the raw and source pages as gzipped CSV into gpurun_out/ and deletes
# the (50+ MB) report, so that the evidence fits gpurun's 64 MiB return limit.
name=$1; regex=$2; skip=$3; count=$4; shift 4
ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c "$count" -o "/tmp/$name" "$@" > "gpurun_out/$name.log" 2>&1
ncu -i "/tmp/$name.ncu-rep" --page raw --csv 2>/dev/null | gzip > "gpurun_out/$name.raw.csv.gz"
ncu -i "/tmp/$name.ncu-rep" --page source --csv 2>/dev/null | gzip > "gpurun_out/$name.source.csv.gz"
python tools/ncu_summary.py "/tmp/$name.ncu-rep" > "gpurun_out/$name.summary.txt" 2>&1
rm -f "/tmp/$name.ncu-rep"
ls -la gpurun_out/$name.*
