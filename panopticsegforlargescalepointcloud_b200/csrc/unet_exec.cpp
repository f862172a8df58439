// Whole-network executor: walks the static tape of the sparse ResUNet and launches the kernels of this library
// (include/pgs_b200.h "Whole-network executor").  Host code only; replaces ~500 Python -> ctypes round trips per
// training step (the step was bound by the host's launch rate: profiles/r1_host_profile.txt).
//
// Reference control flow being replaced: torch_points3d/applications/minkowski.py:160-196 (skip stack),
// modules/MinkowskiEngine/api_modules.py:76-82 (residual block), 281-285 (ResNetDown), 306-311 (ResNetUp), and
// autograd's reverse walk over the same graph.
#include <nvtx3/nvToolsExt.h>   // header-only; ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace {

int conv_launch(int kind, const float* X, const float* W, const int32_t* nbr, const int32_t* order, int64_t n_q, int K,
                int c_in, int c_out, int mirror, int w_transposed, float* Y, void* wprep, size_t wprep_bytes,
                void* stream) {
  switch (kind) {
    case 1: return pgs_conv_fwd_tc(X, nullptr, nbr, order, n_q, K, c_in, c_out, mirror, w_transposed, Y, wprep, wprep_bytes, stream);
    case 2: return pgs_conv_fwd_mma(X, nullptr, nbr, order, n_q, K, c_in, c_out, mirror, w_transposed, Y, wprep, wprep_bytes, stream);
    case 3: return pgs_conv_fwd_mma_split(X, nullptr, nbr, order, n_q, K, c_in, c_out, mirror, w_transposed, Y, wprep, wprep_bytes, stream);
    default: return pgs_conv_fwd(X, W, nbr, n_q, K, c_in, c_out, mirror, w_transposed, Y, stream);
  }
}

// events for the main -> side stream hand-over of each weight gradient (created once per thread, never destroyed)
thread_local std::vector<cudaEvent_t> t_events;

cudaEvent_t event_at(size_t i) {
  while (t_events.size() <= i) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    t_events.push_back(e);
  }
  return t_events[i];
}

}  // namespace

extern "C" {

void pgs_unet_record_bytes(int32_t* out3) {
  out3[0] = (int32_t)sizeof(pgs_unet_op);
  out3[1] = (int32_t)sizeof(pgs_unet_conv);
  out3[2] = (int32_t)sizeof(pgs_unet_bn);
}

int pgs_unet_forward(const pgs_unet_op* ops, int32_t n_ops, int32_t n_slots, float* const* slot_ptr,
                     const int64_t* slot_n, const int32_t* slot_c, const pgs_unet_conv* convs, const pgs_unet_bn* bns,
                     double* sums, float* stats, const int64_t* stat_off, void* stream) {
  PGS_CHECK_ARG(ops && slot_ptr && slot_n && slot_c && n_ops >= 0 && n_slots >= 1, "bad tape");
  struct Range {
    Range(const char* n) { nvtxRangePushA(n); }
    ~Range() { nvtxRangePop(); }
  } range("pgs_unet_forward");
  for (int32_t i = 0; i < n_ops; ++i) {
    const pgs_unet_op& op = ops[i];
    PGS_CHECK_ARG(op.a >= 0 && op.a < n_slots && op.dst > 0 && op.dst < n_slots && op.b < n_slots, "slot out of range");
    const int a = op.a, d = op.dst;
    int rc = PGS_OK;
    switch (op.kind) {
      case PGS_OP_CONV: {
        const pgs_unet_conv& c = convs[op.idx];
        PGS_CHECK_ARG(slot_c[a] == c.c_in && slot_c[d] == c.c_out, "convolution channels do not match its slots");
        rc = conv_launch(c.kind_f, slot_ptr[a], c.W, c.nbr_f, c.order_f, slot_n[d], c.K, c.c_in, c.c_out, c.mirror_f, 0,
                         slot_ptr[d], c.wprep_f, (size_t)c.wprep_bytes, stream);
        break;
      }
      case PGS_OP_BN: {
        const pgs_unet_bn& b = bns[op.idx];
        const int C = slot_c[a];
        const int64_t so = stat_off[op.idx];
        rc = pgs_bn_forward_ex(slot_ptr[a], slot_n[a], C, b.weight, b.bias, b.running_mean, b.running_var, b.training,
                               b.momentum, b.eps, op.relu, PGS_BN_SUMS_ZEROED, sums + so, stats + so, stats + so + C,
                               slot_ptr[d], stream);
        break;
      }
      case PGS_OP_ADD:
        PGS_CHECK_ARG(op.b >= 0, "add needs two inputs");
        rc = pgs_add2(slot_ptr[a], slot_ptr[op.b], slot_ptr[d], slot_n[a] * slot_c[a], stream);
        break;
      case PGS_OP_CAT:
        PGS_CHECK_ARG(op.b >= 0, "cat needs two inputs");
        rc = pgs_cat2(slot_ptr[a], slot_c[a], slot_ptr[op.b], slot_c[op.b], slot_ptr[d], slot_n[a], 0, stream);
        break;
      default:
        PGS_CHECK_ARG(false, "unknown op kind");
    }
    if (rc != PGS_OK) return rc;
  }
  return PGS_OK;
}

int64_t pgs_unet_backward_scratch_elems(const pgs_unet_op* ops, int32_t n_ops, int32_t n_slots, const int64_t* slot_n,
                                        const int32_t* slot_c) {
  // one buffer per conv / bn input gradient, two per cat, one per extra consumer of a multiply-used slot
  std::vector<int> consumers((size_t)n_slots, 0);
  int64_t total = 0;
  for (int32_t i = 0; i < n_ops; ++i) {
    const pgs_unet_op& op = ops[i];
    consumers[op.a]++;
    if (op.b >= 0) consumers[op.b]++;
    if (op.kind == PGS_OP_CONV || op.kind == PGS_OP_BN) total += slot_n[op.a] * slot_c[op.a];
    else if (op.kind == PGS_OP_CAT) total += slot_n[op.a] * (slot_c[op.a] + slot_c[op.b]);
  }
  for (int32_t s = 0; s < n_slots; ++s)
    if (consumers[s] > 1) total += (int64_t)(consumers[s] - 1) * slot_n[s] * slot_c[s];
  return total;
}

int pgs_unet_backward(const pgs_unet_op* ops, int32_t n_ops, int32_t n_slots, int32_t out_slot, float* const* slot_ptr,
                      const int64_t* slot_n, const int32_t* slot_c, const pgs_unet_conv* convs, const pgs_unet_bn* bns,
                      double* sums, const float* stats, const int64_t* stat_off, const float* d_out, float* garena,
                      int64_t garena_elems, float** grad_in, int32_t* n_grad_in, void* stream, void* side_stream) {
  PGS_CHECK_ARG(ops && slot_ptr && slot_n && slot_c && grad_in && n_grad_in, "bad tape");
  PGS_CHECK_ARG(out_slot >= 0 && out_slot < n_slots, "output slot out of range");
  struct Range {
    Range(const char* n) { nvtxRangePushA(n); }
    ~Range() { nvtxRangePop(); }
  } range("pgs_unet_backward");
  cudaStream_t s_main = (cudaStream_t)stream, s_side = (cudaStream_t)side_stream;
  std::vector<std::vector<const float*>> glist((size_t)n_slots);
  glist[out_slot].push_back(d_out);
  int64_t used = 0;
  size_t n_ev = 0;
  auto take = [&](int64_t elems) -> float* {
    float* p = garena + used;
    used += elems;
    return p;
  };
  if (s_side) {  // the side stream may read everything the main stream has produced so far
    cudaEvent_t e = event_at(n_ev++);
    PGS_CHECK_ARG(e != nullptr, "cannot create CUDA event");
    PGS_CUDA(cudaEventRecord(e, s_main));
    PGS_CUDA(cudaStreamWaitEvent(s_side, e, 0));
  }
  for (int32_t i = n_ops - 1; i >= 0; --i) {
    const pgs_unet_op& op = ops[i];
    const int a = op.a, d = op.dst;
    std::vector<const float*>& gl = glist[d];
    if (gl.empty()) continue;
    const int64_t nd = slot_n[d] * slot_c[d];
    const float* g = gl[0];
    for (size_t j = 1; j < gl.size(); ++j) {  // tensor with several consumers: sum their gradients
      PGS_CHECK_ARG(used + nd <= garena_elems, "gradient arena too small");
      float* buf = take(nd);
      int rc = pgs_add2(g, gl[j], buf, nd, stream);
      if (rc) return rc;
      g = buf;
    }
    int rc = PGS_OK;
    switch (op.kind) {
      case PGS_OP_CONV: {
        const pgs_unet_conv& c = convs[op.idx];
        if (c.need_dx) {
          const int64_t na = slot_n[a] * slot_c[a];
          PGS_CHECK_ARG(used + na <= garena_elems, "gradient arena too small");
          float* dx = take(na);
          rc = conv_launch(c.kind_b, g, c.W, c.nbr_b, c.order_b, slot_n[a], c.K, c.c_out, c.c_in, c.mirror_b, 1, dx,
                           c.wprep_b, (size_t)c.wprep_bytes, stream);
          if (rc) return rc;
          glist[a].push_back(dx);
        }
        if (c.dW) {
          void* sw = stream;
          if (s_side) {  // g is complete on the main stream here
            cudaEvent_t e = event_at(n_ev++);
            PGS_CHECK_ARG(e != nullptr, "cannot create CUDA event");
            PGS_CUDA(cudaEventRecord(e, s_main));
            PGS_CUDA(cudaStreamWaitEvent(s_side, e, 0));
            sw = side_stream;
          }
          if (c.pair_in)
            rc = pgs_conv_bwd_weight(slot_ptr[a], g, c.pair_in, c.pair_out, c.pair_offs, c.max_pairs, c.K, c.c_in,
                                     c.c_out, c.mirror_f, c.dW, sw);
          else
            rc = pgs_conv_bwd_weight(slot_ptr[a], g, nullptr, nullptr, nullptr, slot_n[a], 1, c.c_in, c.c_out, 0, c.dW, sw);
        }
        break;
      }
      case PGS_OP_BN: {
        const pgs_unet_bn& b = bns[op.idx];
        const int C = slot_c[a];
        const int64_t na = slot_n[a] * C, so = stat_off[op.idx];
        PGS_CHECK_ARG(used + na <= garena_elems, "gradient arena too small");
        float* dx = take(na);
        const int flags = PGS_BN_SUMS_ZEROED | (b.accumulate ? PGS_BN_ACCUMULATE_PARAM_GRADS : 0);
        // Y = NULL: the ReLU mask is recomputed from x (one array less to read); PGS_BN_MASK=y reads it from Y
        static const bool mask_from_y = [] { const char* e = getenv("PGS_BN_MASK"); return e && e[0] == 'y'; }();
        rc = pgs_bn_backward_ex(slot_ptr[a], (op.relu && mask_from_y) ? slot_ptr[d] : nullptr, g, slot_n[a], C, b.weight,
                                b.bias, stats + so, stats + so + C,
                                b.training, op.relu, flags, sums + so, dx, b.dweight, b.dbias, stream);
        glist[a].push_back(dx);
        break;
      }
      case PGS_OP_ADD:
        glist[a].push_back(g);
        glist[op.b].push_back(g);
        break;
      case PGS_OP_CAT: {
        const int64_t na = slot_n[a] * slot_c[a], nb = slot_n[op.b] * slot_c[op.b];
        PGS_CHECK_ARG(used + na + nb <= garena_elems, "gradient arena too small");
        float* ga = take(na);
        float* gb = take(nb);
        rc = pgs_cat2(ga, slot_c[a], gb, slot_c[op.b], const_cast<float*>(g), slot_n[a], 1, stream);
        glist[a].push_back(ga);
        glist[op.b].push_back(gb);
        break;
      }
      default:
        PGS_CHECK_ARG(false, "unknown op kind");
    }
    if (rc != PGS_OK) return rc;
  }
  if (s_side) {  // gradients complete (and the arenas reusable) for whatever follows on the main stream
    cudaEvent_t e = event_at(n_ev++);
    PGS_CHECK_ARG(e != nullptr, "cannot create CUDA event");
    PGS_CUDA(cudaEventRecord(e, s_side));
    PGS_CUDA(cudaStreamWaitEvent(s_main, e, 0));
  }
  const std::vector<const float*>& g0 = glist[0];
  PGS_CHECK_ARG(g0.size() <= 8, "more than 8 consumers of the network input");
  *n_grad_in = (int32_t)g0.size();
  for (size_t j = 0; j < g0.size(); ++j) grad_in[j] = const_cast<float*>(g0[j]);
  return PGS_OK;
}

}  // extern "C"
