#!/bin/bash
# A/B: variant build (_ab/libsg_old.so) against the working tree, same box
cp spin_glass_anneal_rl_b200/libsg_b200.so /tmp/new.so
for i in 1 2; do
echo "== tree"; timeout 300 python tools/tc_check.py perf 2>&1 | grep "P=3\|P=2"
cp _ab/libsg_old.so spin_glass_anneal_rl_b200/libsg_b200.so
echo "== variant"; SG_TC_VERBOSE=$i timeout 300 python tools/tc_check.py perf 2>&1 | grep "P=3\|P=2\|grid=148" | sort -u
cp /tmp/new.so spin_glass_anneal_rl_b200/libsg_b200.so
done
