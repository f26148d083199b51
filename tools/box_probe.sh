#!/bin/bash
# Records what the GPU box offers for (a) a real GL implementation to run the reference's shaders with, (b) NUMA placement.
out=gpurun_out/box_probe.txt
{
echo "== date"; date -u
echo "== nvidia-smi -L"; nvidia-smi -L
echo "== NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
echo "== GL / EGL / Mesa libraries anywhere"
find / -xdev \( -name 'libOSMesa*' -o -name 'libEGL*' -o -name 'libGL.so*' -o -name 'libGLX*' -o -name 'libGLES*' -o -name '*swrast*' -o -name '*llvmpipe*' -o -name 'libgallium*' -o -name 'libnvidia-egl*' -o -name 'libnvidia-gl*' -o -name 'libglapi*' -o -name 'libvulkan*' -o -name 'libnvidia-glcore*' -o -name 'libGLdispatch*' \) 2>/dev/null | head -100
echo "== mounted driver libs"; ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E 'nvidia|cuda' | head -80
echo "== egl vendor json"; ls /usr/share/glvnd/egl_vendor.d /etc/glvnd/egl_vendor.d 2>/dev/null
echo "== python GL modules"; python - <<'PY'
for m in ("OpenGL", "moderngl", "glfw", "pyglet", "vispy", "pyrender", "vtk", "glcontext"):
    try:
        __import__(m); print(m, "importable")
    except Exception as e:
        print(m, "no:", type(e).__name__)
PY
echo "== Xvfb / glxinfo / eglinfo"; which Xvfb glxinfo eglinfo vulkaninfo 2>&1
echo "== /dev/dri"; ls -la /dev/dri 2>&1
echo "== NUMA"; lscpu | grep -i -E 'numa|socket|model name|^cpu\(s\)'; which numactl; ls /sys/devices/system/node/
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done
nvidia-smi topo -m 2>&1 | head -14
echo "== affinity of this shell"; taskset -p $$; nproc
cat /proc/self/status | grep -i -E 'Mems_allowed_list|Cpus_allowed_list'
echo "== mem"; free -g | head -2
} > $out 2>&1
echo probe done
