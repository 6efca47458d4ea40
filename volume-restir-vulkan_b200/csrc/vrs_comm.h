// Halo exchange between screen-space bands over NCCL (send/recv on NVLink 5 / NVSwitch).  NCCL is bound at run time
// with dlopen so that (a) a Python launcher that already loaded torch's bundled libnccl shares that copy and
// (b) single-GPU users need no NCCL at all.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace vrs {
struct Comm;
bool comm_unique_id(uint8_t id128[128], std::string& err);
Comm* comm_create(const uint8_t id128[128], int rank, int nranks, std::string& err);
void comm_destroy(Comm* c);
// For every plane (RGBA32F, `width` pixels per row, first stored row = store_y0): send the first / last rows of the
// own band [band_y0, band_y1) to the previous / next rank and receive their rows into the halo rows.
bool comm_exchange_halo(Comm* c, cudaStream_t stream, const std::vector<float4*>& planes, uint32_t width, int band_y0, int band_y1,
                        int store_y0, int store_y1, int height, int max_rows, std::string& err);
}  // namespace vrs
