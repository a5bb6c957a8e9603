#!/bin/bash
# usage: cap.sh <tag> <kernel regex> <arch> <nfits> [skip]  -- one ncu --set full capture of a kernel of profiles/prof_fit.py + text summaries
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$2 -s ${5:-2} -c 1 -o $O/$1 -f python profiles/prof_fit.py $3 $4 3 > $O/ncu_$1.log 2>&1
tail -2 $O/ncu_$1.log
python profiles/ncu_summary.py $O/$1.ncu-rep "$1: $3 x $4 fits (prof_fit.py), ncu --set full --clock-control none" > $O/ncu_$1.txt 2>&1
python profiles/ncu_source_lines.py $O/$1.ncu-rep 70 > $O/ncu_$1_lines.txt 2>&1
