/* me_rt.h — run-time compilation / driver-API helpers shared by the translation units of libme_b200.so (implemented in
 * me_api.cu, which owns the lazily loaded NVRTC and driver entry points).  Internal: not part of the C ABI. */
#ifndef ME_RT_H
#define ME_RT_H

#include <cuda.h>

#include <string>
#include <vector>

/* NVRTC-compile `src` for sm_100a with the library's kernel headers available as in-memory includes; opts are extra
 * options ("-DNAME=value").  Returns an me_status_code; `log` receives the compiler log. */
int me_rt_compile(const std::string &src, const char *name, const std::vector<std::string> &opts, std::vector<char> &cubin,
                  std::string &log);
/* Load a CUBIN on `device` and look up `n` kernels by name. */
int me_rt_load(int device, const std::vector<char> &cubin, const char *const *names, int n, CUfunction *out, std::string &err);
int me_rt_launch(CUfunction f, unsigned grid, unsigned block, unsigned smem, void *stream, void **args, std::string &err);
int me_rt_set_dynamic_smem(CUfunction f, int bytes, std::string &err);
/* 2-D tensor map over a row-major BF16 array [rows][inner] (no swizzle, no interleave) with box [box_rows][box_inner];
 * map_out = 128 bytes, 64-byte aligned (a CUtensorMap).  ME_ERR_UNSUPPORTED when the driver has no tensor maps. */
int me_rt_tensor_map_2d_bf16(void *map_out, const void *gaddr, unsigned long long inner, unsigned long long rows,
                             unsigned box_inner, unsigned box_rows);

#endif
