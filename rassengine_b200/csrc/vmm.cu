// Growable device arrays on the CUDA virtual-memory API: the store's big arrays (fp32 rows, bf16 shadow, norms, scan
// scale / offset) reserve their address range once and map physical memory in chunks as rows arrive.
//
// Growing by cudaMalloc + copy + cudaFree needs the old and the new array resident at once: a 102 GB bf16 shard (half
// of BASELINE configs[4] on two GPUs) could never grow, and a doubling policy wastes up to half of HBM.  Here growth
// maps more chunks behind the same base pointer: no copy, no second resident array, pointers (and the TMA tensor map's
// base) stay valid.  The driver entry points are resolved at run time (the library does not link libcuda); if any is
// missing or a call fails before the first mapping, the engine falls back to cudaMalloc + copy.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace {

struct VmApi {
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

VmApi vm_load() {
  VmApi api;
  auto get = [](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess &&
           *fn != nullptr;
  };
  api.ok = get("cuMemAddressReserve", (void**)&api.reserve) && get("cuMemAddressFree", (void**)&api.addr_free) &&
           get("cuMemCreate", (void**)&api.create) && get("cuMemRelease", (void**)&api.release) &&
           get("cuMemMap", (void**)&api.map) && get("cuMemUnmap", (void**)&api.unmap) &&
           get("cuMemSetAccess", (void**)&api.set_access) &&
           get("cuMemGetAllocationGranularity", (void**)&api.granularity);
  if (getenv("RASS_DEBUG_NO_VMM")) api.ok = false;          // A/B and test switch: the cudaMalloc + copy path
  return api;
}

// (a function-local static: initialised once, also when the shards of one handle grow from their worker threads)
const VmApi& vm_api() {
  static const VmApi api = vm_load();
  return api;
}

CUmemAllocationProp vm_prop(int device) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}

}  // namespace

bool vm_available() { return vm_api().ok; }

// Reserves `max_bytes` of address space (rounded up to the chunk size) for one array.  False: use cudaMalloc.
bool vm_reserve(VmArray* a, int device, size_t max_bytes, size_t chunk_hint) {
  const VmApi& api = vm_api();
  if (!api.ok || a->base) return false;
  const CUmemAllocationProp prop = vm_prop(device);
  size_t gran = 0;
  if (api.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return false;
  // chunks of `chunk_hint` bytes (32k rows of the array), at most 256 MB, a multiple of the granularity: a few
  // thousand mappings for a 100 GB array, a few hundred MB of slack for a small engine
  size_t chunk = std::min<size_t>(std::max<size_t>(gran, chunk_hint), (size_t)256 << 20);
  chunk = (chunk + gran - 1) / gran * gran;
  const size_t reserved = (max_bytes + chunk - 1) / chunk * chunk;
  CUdeviceptr base = 0;
  if (api.reserve(&base, reserved, 0, 0, 0) != CUDA_SUCCESS) return false;
  a->base = (void*)base;
  a->reserved = reserved;
  a->mapped = 0;
  a->chunk = chunk;
  a->device = device;
  return true;
}

// Makes the first `bytes` of the array usable.  0 = ok, 1 = out of memory, 2 = any other failure.
int vm_grow(VmArray* a, size_t bytes) {
  const VmApi& api = vm_api();
  if (bytes <= a->mapped) return 0;
  if (bytes > a->reserved) return 2;
  const CUmemAllocationProp prop = vm_prop(a->device);
  CUmemAccessDesc acc = {};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  while (a->mapped < bytes) {
    CUmemGenericAllocationHandle hd;
    CUresult r = api.create(&hd, a->chunk, &prop, 0);
    if (r != CUDA_SUCCESS) return r == CUDA_ERROR_OUT_OF_MEMORY ? 1 : 2;
    const CUdeviceptr at = (CUdeviceptr)a->base + a->mapped;
    r = api.map(at, a->chunk, 0, hd, 0);
    if (r == CUDA_SUCCESS) {
      r = api.set_access(at, a->chunk, &acc, 1);
      if (r != CUDA_SUCCESS) api.unmap(at, a->chunk);
    }
    if (r != CUDA_SUCCESS) {
      api.release(hd);
      return r == CUDA_ERROR_OUT_OF_MEMORY ? 1 : 2;
    }
    a->handles.push_back((unsigned long long)hd);
    a->mapped += a->chunk;
  }
  return 0;
}

void vm_free(VmArray* a) {
  if (!a->base) return;
  const VmApi& api = vm_api();
  if (a->mapped) api.unmap((CUdeviceptr)a->base, a->mapped);
  for (unsigned long long hd : a->handles) api.release((CUmemGenericAllocationHandle)hd);
  api.addr_free((CUdeviceptr)a->base, a->reserved);
  a->handles.clear();
  a->base = nullptr;
  a->reserved = a->mapped = 0;
}
