// Stand-alone SENetLayer / BilinearInteraction entry points (module-level API of the reference,
// src/model_fibinet.py:5-89) for arbitrary field counts.
#include "common.cuh"
#include "gemm.h"
#include "tower.h"

using namespace fbn;

extern "C" size_t fbn_senet_scratch_bytes(int64_t batch, int fields, int hidden) {
  return (size_t)batch * (2 * fields + 2 * hidden) * sizeof(float) + (size_t)1024 * (2 * fields * hidden + fields + hidden) * sizeof(float);
}
extern "C" size_t fbn_bilinear_scratch_bytes(int64_t batch, int fields, int dim, int type) {
  const int64_t P = (int64_t)fields * (fields - 1) / 2;
  const int64_t nT = type == FBN_BILINEAR_ALL ? fields : (type == FBN_BILINEAR_EACH ? fields - 1 : P);
  const int64_t nW = type == FBN_BILINEAR_ALL ? 1 : nT;
  return (size_t)(2 * batch * nT * dim + 32 * nW * dim * dim) * sizeof(float);
}
extern "C" int fbn_senet_fwd(const float*, const float*, const float*, const float*, const float*, int64_t, int, int, int, float*,
                             float*, fbn_stream_t) {
  set_error("fbn_senet_fwd: not implemented yet");
  return FBN_ERR_ARG;
}
extern "C" int fbn_senet_bwd(const float*, const float*, const float*, const float*, const float*, const float*, int64_t, int, int, int,
                             float*, float*, float*, float*, float*, void*, size_t, fbn_stream_t) {
  set_error("fbn_senet_bwd: not implemented yet");
  return FBN_ERR_ARG;
}
extern "C" int fbn_bilinear_fwd(const float*, const float*, int, int64_t, int, int, float*, void*, size_t, int, fbn_stream_t) {
  set_error("fbn_bilinear_fwd: not implemented yet");
  return FBN_ERR_ARG;
}
extern "C" int fbn_bilinear_bwd(const float*, const float*, const float*, int, int64_t, int, int, float*, float*, void*, size_t, int,
                                fbn_stream_t) {
  set_error("fbn_bilinear_bwd: not implemented yet");
  return FBN_ERR_ARG;
}
