// gm_peer.cu -- multi-GPU exchange over NVLink peer memory (include/geosmie_b200.h, "multi-GPU exchange").
//
// One process per GPU.  Rank 0 owns the gather buffer (cudaMalloc + cudaIpcGetMemHandle); the other ranks map it with
// cudaIpcOpenMemHandle (which enables peer access lazily) and write their rows into their own segment, either with
// copy-engine transfers (gm_peer_put: no SM time on either GPU, overlaps the next step's kernels) or straight from the
// producing kernels (gm_table_set_mirror: P2P stores).  There is no reference counterpart: the reference builds a table in
// one Python process (dointegration.py:804-810 notes the cells are independent).
#include <string.h>

#include "gm_common.cuh"

static int peer_streams(gm_handle_s* h) {
  if (h->peer_stream) return GM_OK;
  GM_CUDA_TRY(cudaStreamCreateWithFlags(&h->peer_stream, cudaStreamNonBlocking));
  GM_CUDA_TRY(cudaEventCreateWithFlags(&h->peer_ev_compute, cudaEventDisableTiming));
  GM_CUDA_TRY(cudaEventCreateWithFlags(&h->peer_ev_done, cudaEventDisableTiming));
  return GM_OK;
}

extern "C" int gm_peer_alloc(gm_handle_t h, size_t bytes, void** dptr, unsigned char ipc_handle[GM_IPC_HANDLE_BYTES]) {
  GM_REQUIRE(h != nullptr && dptr != nullptr && ipc_handle != nullptr, "NULL argument");
  GM_REQUIRE(bytes > 0, "empty exchange buffer");
  static_assert(sizeof(cudaIpcMemHandle_t) == GM_IPC_HANDLE_BYTES, "IPC handle size");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    gm_set_error("cudaMalloc(%zu) for the exchange buffer failed: %s", bytes, cudaGetErrorString(e));
    return GM_ENOMEM;
  }
  cudaIpcMemHandle_t hd;
  e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    gm_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return GM_ECUDA;
  }
  memcpy(ipc_handle, &hd, GM_IPC_HANDLE_BYTES);
  *dptr = p;
  return GM_OK;
}

extern "C" int gm_peer_free(gm_handle_t h, void* dptr) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  if (!dptr) return GM_OK;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  GM_CUDA_TRY(cudaFree(dptr));
  return GM_OK;
}

extern "C" int gm_peer_open(gm_handle_t h, const unsigned char ipc_handle[GM_IPC_HANDLE_BYTES], void** dptr) {
  GM_REQUIRE(h != nullptr && dptr != nullptr && ipc_handle != nullptr, "NULL argument");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, ipc_handle, GM_IPC_HANDLE_BYTES);
  void* p = nullptr;
  GM_CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *dptr = p;
  return GM_OK;
}

extern "C" int gm_peer_close(gm_handle_t h, void* dptr) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  if (!dptr) return GM_OK;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  GM_CUDA_TRY(cudaIpcCloseMemHandle(dptr));
  return GM_OK;
}

extern "C" int gm_peer_put(gm_handle_t h, void* dst, const void* src, size_t bytes) {
  GM_REQUIRE(h != nullptr && dst != nullptr && src != nullptr, "NULL argument");
  if (bytes == 0) return GM_OK;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  int rc = peer_streams(h);
  if (rc) return rc;
  // ordered after the kernels enqueued so far on the compute stream; later kernels are free to overlap the transfer
  GM_CUDA_TRY(cudaEventRecord(h->peer_ev_compute, h->stream));
  GM_CUDA_TRY(cudaStreamWaitEvent(h->peer_stream, h->peer_ev_compute, 0));
  GM_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->peer_stream));
  h->peer_pending = true;
  return GM_OK;
}

// Strided read-out on the exchange stream (ordered after the compute stream): `height` rows of `width` bytes from device / peer
// memory into host memory with a row pitch -- rank 0 scatters the segment of rank r into rows r, r + W, ... of the final table arrays.
extern "C" int gm_peer_get2d(gm_handle_t h, void* dst_host, size_t dpitch, const void* src_dev, size_t spitch, size_t width, size_t height) {
  GM_REQUIRE(h != nullptr && dst_host != nullptr && src_dev != nullptr, "NULL argument");
  GM_REQUIRE(dpitch >= width && spitch >= width, "pitch smaller than the row width");
  if (width == 0 || height == 0) return GM_OK;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  int rc = peer_streams(h);
  if (rc) return rc;
  GM_CUDA_TRY(cudaEventRecord(h->peer_ev_compute, h->stream));
  GM_CUDA_TRY(cudaStreamWaitEvent(h->peer_stream, h->peer_ev_compute, 0));
  GM_CUDA_TRY(cudaMemcpy2DAsync(dst_host, dpitch, src_dev, spitch, width, height, cudaMemcpyDeviceToHost, h->peer_stream));
  h->peer_pending = true;
  return GM_OK;
}

extern "C" int gm_peer_join(gm_handle_t h) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  if (!h->peer_pending) return GM_OK;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  GM_CUDA_TRY(cudaEventRecord(h->peer_ev_done, h->peer_stream));
  GM_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->peer_ev_done, 0));
  return GM_OK;
}

extern "C" int gm_peer_sync(gm_handle_t h) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  if (!h->peer_stream) return GM_OK;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  GM_CUDA_TRY(cudaStreamSynchronize(h->peer_stream));
  h->peer_pending = false;
  return GM_OK;
}

// Marks: gm_peer_mark(i) records "every put issued so far" on the exchange stream; gm_peer_wait(i) makes the compute stream
// wait for that point -- the fence a double-buffered producer needs before it overwrites the source of an earlier put,
// without giving up the overlap of the most recent one.  A mark that was never recorded does not block.
extern "C" int gm_peer_mark(gm_handle_t h, int idx) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  GM_REQUIRE(idx >= 0 && idx < 4, "mark index out of range (0..3)");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  int rc = peer_streams(h);
  if (rc) return rc;
  if (!h->peer_marks[idx]) GM_CUDA_TRY(cudaEventCreateWithFlags(&h->peer_marks[idx], cudaEventDisableTiming));
  GM_CUDA_TRY(cudaEventRecord(h->peer_marks[idx], h->peer_stream));
  return GM_OK;
}

extern "C" int gm_peer_wait(gm_handle_t h, int idx) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  GM_REQUIRE(idx >= 0 && idx < 4, "mark index out of range (0..3)");
  if (!h->peer_marks[idx]) return GM_OK;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  GM_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->peer_marks[idx], 0));
  return GM_OK;
}
