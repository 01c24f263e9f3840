#!/bin/bash
# A/B timing: current library vs dietgpu_b200/libdietgpu_b200_old.so
cd /root/repo
for wl in "$@"; do python tools/sweep.py $wl "parts=1" 2>&1 | tail -1; done
cp dietgpu_b200/libdietgpu_b200.so /tmp/new.so; cp dietgpu_b200/libdietgpu_b200_old.so dietgpu_b200/libdietgpu_b200.so
echo "--- old"
for wl in "$@"; do python tools/sweep.py $wl "parts=1" 2>&1 | tail -1; done
cp /tmp/new.so dietgpu_b200/libdietgpu_b200.so
