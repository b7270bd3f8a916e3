#!/bin/bash
# round 2, call D: stage-1 GL bring-up through the NVIDIA vendor library (oracle/glsl/egl_direct.c)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( cat /proc/driver/nvidia/version; nvidia-smi --query-gpu=driver_version --format=csv; ls -la /usr/local/nvidia/lib64/ | grep -E "libcuda|libnvidia-ml|EGL|glsi|rtcore|nvidia-gpucomp" ; ls /usr/lib/x86_64-linux-gnu | grep -E "libcuda|libnvidia" | head; env | grep -i nvidia ) > $O/r2d_versions.txt 2>&1
timeout 60 strace -f -e trace=openat,ioctl -o $O/r2d_strace.txt ./oracle/_ref/gl_probe > $O/r2d_gl_probe.txt 2>&1; echo "rc=$?" >> $O/r2d_gl_probe.txt
which strace >> $O/r2d_versions.txt 2>&1
timeout 60 ./oracle/_ref/gl_probe >> $O/r2d_gl_probe.txt 2>&1; echo "rc=$?" >> $O/r2d_gl_probe.txt
cat $O/r2d_versions.txt; cat $O/r2d_gl_probe.txt; grep -E "nvidia|dri" $O/r2d_strace.txt | tail -40
