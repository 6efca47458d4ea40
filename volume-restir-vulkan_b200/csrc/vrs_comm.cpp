#include "vrs_comm.h"

#include <dlfcn.h>

#include <cstring>

namespace vrs {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat = 7;   // ncclFloat32

struct Api {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Api* api(std::string& err) {
  static Api a;
  if (a.handle) return &a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (a.handle) break; }
  if (!a.handle) for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
  if (!a.handle) { err = std::string("NCCL not found: ") + dlerror(); return nullptr; }
  bool ok = true;
  auto sym = [&](const char* s) { void* p = dlsym(a.handle, s); if (!p) ok = false; return p; };
  a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
  a.Send = (decltype(a.Send))sym("ncclSend");
  a.Recv = (decltype(a.Recv))sym("ncclRecv");
  a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
  a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
  if (!ok) { err = "NCCL library lacks required symbols"; a.handle = nullptr; return nullptr; }
  return &a;
}
}  // namespace

struct Comm { ncclComm_t comm = nullptr; int rank = 0, nranks = 1; };

bool comm_unique_id(uint8_t id128[128], std::string& err) {
  Api* a = api(err); if (!a) return false;
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != 0) { err = std::string("ncclGetUniqueId: ") + a->GetErrorString(r); return false; }
  memcpy(id128, id.internal, 128);
  return true;
}

Comm* comm_create(const uint8_t id128[128], int rank, int nranks, std::string& err) {
  Api* a = api(err); if (!a) return nullptr;
  ncclUniqueId id; memcpy(id.internal, id128, 128);
  Comm* c = new Comm(); c->rank = rank; c->nranks = nranks;
  ncclResult_t r = a->CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) { err = std::string("ncclCommInitRank: ") + a->GetErrorString(r); delete c; return nullptr; }
  return c;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  std::string err; Api* a = api(err);
  if (a && c->comm) a->CommDestroy(c->comm);
  delete c;
}

bool comm_exchange_halo(Comm* c, cudaStream_t stream, const std::vector<float4*>& planes, uint32_t width, int band_y0, int band_y1,
                        int store_y0, int store_y1, int height, int max_rows, std::string& err) {
  Api* a = api(err); if (!a) return false;
  const int up = c->rank - 1, down = c->rank + 1;
  int halo_up = band_y0 - store_y0, halo_down = store_y1 - band_y1;           // rows we receive
  // rows the neighbours keep of our band = their halo; bands are uniform so it equals our own configured halo,
  // clipped by the image: the previous rank stores [its band_y1, its band_y1 + halo) = our first rows.
  const int cfg_halo = halo_up > halo_down ? halo_up : halo_down;
  int send_up = cfg_halo, send_down = cfg_halo;
  if (send_up > band_y1 - band_y0) send_up = band_y1 - band_y0;
  if (send_down > band_y1 - band_y0) send_down = band_y1 - band_y0;
  if (send_up > max_rows) send_up = max_rows;                                  // (spatial exchanges move ceil(radius) rows, not the whole halo)
  if (send_down > max_rows) send_down = max_rows;
  if (halo_up > max_rows) halo_up = max_rows;
  if (halo_down > max_rows) halo_down = max_rows;
  (void)height;
  auto row_ptr = [&](float4* p, int row) { return p + (size_t)(row - store_y0) * width; };
  ncclResult_t r = a->GroupStart();
  for (float4* p : planes) {
    if (r != 0) break;
    if (up >= 0) {
      if (send_up > 0) r = a->Send(row_ptr(p, band_y0), (size_t)send_up * width * 4, kNcclFloat, up, c->comm, stream);
      if (r == 0 && halo_up > 0) r = a->Recv(row_ptr(p, band_y0 - halo_up), (size_t)halo_up * width * 4, kNcclFloat, up, c->comm, stream);
    }
    if (r == 0 && down < c->nranks) {
      if (send_down > 0) r = a->Send(row_ptr(p, band_y1 - send_down), (size_t)send_down * width * 4, kNcclFloat, down, c->comm, stream);
      if (r == 0 && halo_down > 0) r = a->Recv(row_ptr(p, band_y1), (size_t)halo_down * width * 4, kNcclFloat, down, c->comm, stream);
    }
  }
  ncclResult_t e = a->GroupEnd();
  if (r == 0) r = e;
  if (r != 0) { err = std::string("halo exchange: ") + a->GetErrorString(r); return false; }
  return true;
}

}  // namespace vrs
