// Host -> device upload of PAGEABLE host vectors at (close to) PCIe rate.
//
// The reference's callers hand MGCMTSolver.vcycle plain numpy arrays (2DPotGS.py:93-95), i.e. pageable memory.  A
// cudaMemcpy from pageable memory is staged by the driver through one internal bounce buffer by ONE host thread
// (~10 GB/s here): for a 4096^2 vector that is 11 ms of upload in front of a 0.5 ms V-cycle.  Here the staging copy is
// done by several host threads into page-locked chunks of our own, each chunk sent with its own asynchronous DMA as soon
// as it is filled, so the host-side copy and the PCIe transfer overlap and the host copy runs at memory bandwidth.
//
// Layout: T worker threads (spawned per call: 8 thread starts per 134 MB vector are noise), each owning two page-locked
// chunks (double buffering: fill one while the other's DMA is in flight) and one event per chunk.  Thread t takes chunks
// t, t + T, t + 2T, ... of the source.  All DMAs go to the caller's stream; the call returns once every chunk has been
// enqueued (the caller records its own event behind them).  The pool of chunks is allocated once and kept.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "kernels.h"

namespace mgcmt {

int g_stage_threads = 8;        // mgcmt_set_option("stage_threads", n): host threads of the staged upload (0: plain cudaMemcpyAsync)
int g_stage_chunk_kib = 4096;   // mgcmt_set_option("stage_chunk_kib", n): bytes per page-locked chunk / 1024

namespace {

struct StagePool {
  std::mutex mu;   // one staged upload at a time
  int device = -1;
  int nthreads = 0;
  size_t chunk = 0;
  std::vector<void *> buf;       // 2 per thread
  std::vector<cudaEvent_t> ev;   // 2 per thread: the DMA out of the chunk has completed
};

StagePool *pool() {
  static StagePool *p = new StagePool();   // never freed: no teardown-order problems with the CUDA runtime at exit
  return p;
}

void release(StagePool &P) {
  for (void *b : P.buf) cudaFreeHost(b);
  for (cudaEvent_t e : P.ev) cudaEventDestroy(e);
  P.buf.clear();
  P.ev.clear();
  P.nthreads = 0;
  P.chunk = 0;
}

cudaError_t prepare(StagePool &P, int device, int nthreads, size_t chunk) {
  if (P.device == device && P.nthreads == nthreads && P.chunk == chunk) return cudaSuccess;
  release(P);
  P.device = device;
  for (int i = 0; i < 2 * nthreads; ++i) {
    void *b = nullptr;
    cudaError_t e = cudaHostAlloc(&b, chunk, cudaHostAllocDefault);
    if (e != cudaSuccess) { release(P); return e; }
    P.buf.push_back(b);
    cudaEvent_t ev;
    e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { release(P); return e; }
    P.ev.push_back(ev);
  }
  P.nthreads = nthreads;
  P.chunk = chunk;
  return cudaSuccess;
}

}  // namespace

bool host_pointer_is_pinned(const void *p) {
  cudaPointerAttributes a;
  const cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();   // plain malloc'ed memory is "invalid value" on older drivers: not an error here
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

cudaError_t staged_upload(void *d_dst, const void *h_src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return cudaSuccess;
  const size_t chunk = (size_t)std::max(64, g_stage_chunk_kib) * 1024;
  const size_t nchunks = (bytes + chunk - 1) / chunk;
  int nthreads = std::min<size_t>((size_t)std::max(0, g_stage_threads), nchunks);
  const unsigned hw = std::thread::hardware_concurrency();
  if (hw > 0) nthreads = std::min<int>(nthreads, (int)hw);
  if (nthreads < 1 || bytes < 4 * chunk || host_pointer_is_pinned(h_src))
    return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream);

  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  StagePool &P = *pool();
  std::lock_guard<std::mutex> lock(P.mu);
  e = prepare(P, device, nthreads, chunk);
  if (e != cudaSuccess) return e;

  std::vector<cudaError_t> err(nthreads, cudaSuccess);
  auto work = [&](int t) {
    cudaError_t le = cudaSetDevice(device);   // the current device is per host thread
    int slot = 0;
    for (size_t j = (size_t)t; j < nchunks && le == cudaSuccess; j += (size_t)nthreads, slot ^= 1) {
      const size_t off = j * chunk, len = std::min(chunk, bytes - off);
      void *b = P.buf[2 * t + slot];
      le = cudaEventSynchronize(P.ev[2 * t + slot]);   // the previous DMA out of this chunk (a never-recorded event is complete)
      if (le != cudaSuccess) break;
      std::memcpy(b, (const char *)h_src + off, len);
      le = cudaMemcpyAsync((char *)d_dst + off, b, len, cudaMemcpyHostToDevice, stream);
      if (le != cudaSuccess) break;
      le = cudaEventRecord(P.ev[2 * t + slot], stream);
    }
    err[t] = le;
  };
  std::vector<std::thread> th;
  th.reserve(nthreads - 1);
  int started = 1;   // thread 0 is the caller
  try {
    for (int t = 1; t < nthreads; ++t) {
      th.emplace_back(work, t);
      ++started;
    }
  } catch (...) {
    // the process may start no more threads: the caller takes over the chunks of the workers that never started
    // (no exception may cross the C ABI)
  }
  work(0);
  for (int t = started; t < nthreads; ++t) work(t);
  for (auto &x : th) x.join();
  for (cudaError_t le : err)
    if (le != cudaSuccess) return le;
  return cudaSuccess;
}

}  // namespace mgcmt
