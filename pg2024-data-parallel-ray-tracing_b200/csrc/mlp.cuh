// mlp.cuh -- interface of the fused proxy-MLP inference kernel (mlp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include "dprt_types.h"

namespace dprt {

struct MlpModel;   // device-resident, pre-tiled weights of one NeuralVisNetworkWith{4,6}Res256SingleOutput

// blob: packed fp32 weights (DESIGN.md "proxy weight blob"); dtype 0 = bf16 operands, 1 = fp16 operands.
int  mlp_create(const void* blob, size_t bytes, int dtype, MlpModel** out, std::string& err);
void mlp_destroy(MlpModel* m);
// x_dev: [n,5] fp16 features, y_dev: [n] fp16 predictions (both device). Asynchronous on `stream`.
int  mlp_forward(const MlpModel* m, const dprt_half* x_dev, dprt_half* y_dev, int64_t n, cudaStream_t stream,
                 std::string& err);
int64_t mlp_macs_per_row(const MlpModel* m);

}  // namespace dprt
