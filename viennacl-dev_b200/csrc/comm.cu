// comm.cu -- communicator plumbing for the row-partitioned (one process per GPU) path.  NCCL is dlopen'ed on first use.
// There is no counterpart in the reference (doc/manual/multi-device.dox:9: "Partition of data is left to the user").
#include "common.cuh"
#include "nccl_dyn.cuh"
#include <dlfcn.h>
#include <cstdlib>

static NcclApi g_api;
static bool g_tried = false, g_ok = false;
static std::string g_why;

const NcclApi *vcl_nccl(const char **why)
{
  if (!g_tried)
  {
    g_tried = true;
    // An already-loaded libnccl (e.g. the one bundled with torch) wins: RTLD_NOLOAD first.
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }
    if (!h)
    {
      const char *env = getenv("VCL_B200_NCCL_LIB");
      if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { g_why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); }
    else
    {
      bool ok = true;
#define VCL_SYM(field, name) do { *(void**)(&g_api.field) = dlsym(h, name); if (!g_api.field) { ok = false; g_why = std::string("missing symbol ") + name; } } while (0)
      VCL_SYM(GetUniqueId, "ncclGetUniqueId");
      VCL_SYM(CommInitRank, "ncclCommInitRank");
      VCL_SYM(CommDestroy, "ncclCommDestroy");
      VCL_SYM(AllReduce, "ncclAllReduce");
      VCL_SYM(AllGather, "ncclAllGather");
      VCL_SYM(Send, "ncclSend");
      VCL_SYM(Recv, "ncclRecv");
      VCL_SYM(GroupStart, "ncclGroupStart");
      VCL_SYM(GroupEnd, "ncclGroupEnd");
      VCL_SYM(GetErrorString, "ncclGetErrorString");
      VCL_SYM(CommGetAsyncError, "ncclCommGetAsyncError");
#undef VCL_SYM
      g_ok = ok;
    }
  }
  if (why) *why = g_why.c_str();
  return g_ok ? &g_api : nullptr;
}

// Poll the communicator for asynchronous errors (ncclCommGetAsyncError): called by the NCCL transport once per solver batch, after
// the stream synchronisation, and by ViennaCLBackendCommCheck.
ViennaCLStatus vcl_comm_check(ViennaCLBackend b)
{
  if (!b->nccl_comm) return ViennaCLSuccess;
  const NcclApi *api = vcl_nccl(nullptr);
  if (!api) return ViennaCLSuccess;
  ncclResult_t async = ncclSuccess;
  const ncclResult_t r = api->CommGetAsyncError((ncclComm_t)b->nccl_comm, &async);
  if (r != ncclSuccess) return vcl_fail(b, ViennaCLB200CommError, api->GetErrorString(r), __FILE__, __LINE__);
  if (async != ncclSuccess && async != ncclInProgress) return vcl_fail(b, ViennaCLB200CommError, api->GetErrorString(async), __FILE__, __LINE__);
  return ViennaCLSuccess;
}

extern "C" {

ViennaCLStatus ViennaCLBackendCommCheck(ViennaCLBackend b)
{
  VCL_CHECK_BACKEND(b);
  return vcl_comm_check(b);
}

ViennaCLStatus ViennaCLBackendCommGetUniqueId(ViennaCLBackend b, void *id_bytes)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, id_bytes, "null id buffer");
  const char *why = nullptr;
  const NcclApi *api = vcl_nccl(&why);
  if (!api) return vcl_fail(b, ViennaCLB200CommError, why, __FILE__, __LINE__);
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return vcl_fail(b, ViennaCLB200CommError, api->GetErrorString(r), __FILE__, __LINE__);
  static_assert(sizeof(ncclUniqueId) == VCL_B200_COMM_ID_BYTES, "id size");
  std::memcpy(id_bytes, &id, sizeof(id));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendCommInit(ViennaCLBackend b, const void *id_bytes, ViennaCLInt rank, ViennaCLInt world_size)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, id_bytes && world_size >= 1 && rank >= 0 && rank < world_size, "bad communicator arguments");
  if (b->nccl_comm) VCL_TRY(ViennaCLBackendCommDestroy(b));
  b->rank = rank; b->world = world_size;
  if (world_size == 1) return ViennaCLSuccess;
  const char *why = nullptr;
  const NcclApi *api = vcl_nccl(&why);
  if (!api) return vcl_fail(b, ViennaCLB200CommError, why, __FILE__, __LINE__);
  VCL_CUDA(b, cudaSetDevice(b->device));
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  ncclResult_t r = api->CommInitRank(&comm, world_size, id, rank);
  if (r != ncclSuccess) return vcl_fail(b, ViennaCLB200CommError, api->GetErrorString(r), __FILE__, __LINE__);
  b->nccl_comm = comm;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendCommDestroy(ViennaCLBackend b)
{
  VCL_CHECK_BACKEND(b);
  if (b->nccl_comm)
  {
    const NcclApi *api = vcl_nccl(nullptr);
    if (api) api->CommDestroy((ncclComm_t)b->nccl_comm);
    b->nccl_comm = nullptr;
  }
  b->rank = 0; b->world = 1;
  return ViennaCLSuccess;
}

} // extern "C"
