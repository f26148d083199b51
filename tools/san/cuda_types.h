// g++ (non-CUDA) build of the host-only sources: the two CUDA vector types internal.h mentions in kernel-launcher signatures
#ifndef __CUDACC__
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
#endif
