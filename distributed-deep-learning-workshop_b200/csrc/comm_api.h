// Plain launch interface of the fused all-reduce kernels (symmetric memory over NVLink 5 / NVSwitch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum CommDtype : int { kF32 = 0, kBF16 = 1 };

struct CommCtx {
  void* const* peer_bufs;       // device array [world]: base of every rank's symmetric data buffer
  uint32_t* const* peer_flags;  // device array [world]: base of every rank's symmetric flag buffer
  void* mc_buf;                 // multicast (NVLS) address of the data buffer, or nullptr
  int rank;
  int world;
};

// out[i] = scale * sum_r in_r[i]  for element range [off, off+n) of the symmetric buffer.
// one-shot: every rank pulls all peers' data (latency path);  dst may be any local buffer (cast fused).
void allreduce_oneshot(const CommCtx& c, int64_t off_elems, int64_t n, CommDtype in_t, void* dst, CommDtype out_t,
                       float scale, int blocks, cudaStream_t s);
// two-shot over P2P loads/stores: reduce my 1/W slice from all peers, scale, write it back into every peer.
void allreduce_twoshot_p2p(const CommCtx& c, int64_t off_elems, int64_t n, CommDtype t, float scale, int blocks,
                           cudaStream_t s);
// two-shot through the switch: multimem.ld_reduce (in-switch reduction) + scale + multimem.st (multicast), in place.
void allreduce_twoshot_nvls(const CommCtx& c, int64_t off_elems, int64_t n, CommDtype t, float scale, int blocks,
                            cudaStream_t s);
// broadcast range from `root` to every rank's symmetric buffer (multicast store if available, else P2P stores).
void broadcast_sym(const CommCtx& c, int64_t off_elems, int64_t n, CommDtype t, int root, int blocks, cudaStream_t s);
// all-reduce fused with the SGD-momentum update: each rank reduces its 1/W slice of the fp32 gradient in the
// switch, updates ITS slice of the fp32 master weights + momentum, and multicasts the new weights (fp32 in the
// symmetric weight buffer and optionally bf16) to all ranks.  hyper layout as in ops_api.h.
void allreduce_sgd_nvls(const CommCtx& grad, const CommCtx& weight, int64_t off_elems, int64_t n, float* mom_local,
                        void* w16_mc, float scale, const float* hyper, int blocks, cudaStream_t s);

constexpr int kCommMaxBlocks = 64;

}  // namespace b200
