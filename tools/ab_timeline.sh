#!/bin/bash
cp spin_glass_anneal_rl_b200/libsg_b200.so /tmp/new.so
echo "== new"; timeout 300 python tools/tc_timeline.py 3 2>&1 | tail -7
cp _ab/libsg_old.so spin_glass_anneal_rl_b200/libsg_b200.so
echo "== old"; timeout 300 python tools/tc_timeline.py 3 2>&1 | tail -7
cp /tmp/new.so spin_glass_anneal_rl_b200/libsg_b200.so
