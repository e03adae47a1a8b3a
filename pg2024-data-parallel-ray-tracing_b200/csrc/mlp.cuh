// mlp.cuh -- interface of the fused proxy-MLP inference kernel (mlp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include "dprt_types.h"

namespace dprt {

// one row of the grouped launch's table: the proxy network of one scene object (wstages == nullptr: no model, rows keep 0)
struct MlpGroupEntry { const uint8_t* wstages; const float* small; int32_t nres; int32_t head; };     // head: 0 = LeakyReLU output, 1 = Sigmoid

struct MlpModel;   // device-resident, pre-tiled weights of one NeuralVisNetworkWith{4,6}Res256SingleOutput

// blob: packed fp32 weights (DESIGN.md "proxy weight blob"); dtype 0 = bf16 operands, 1 = fp16 operands.
int  mlp_create(const void* blob, size_t bytes, int dtype, MlpModel** out, std::string& err);
void mlp_destroy(MlpModel* m);
// x_dev: [n,5] fp16 features, y_dev: [n] fp16 predictions (both device). Asynchronous on `stream`.
int  mlp_forward(const MlpModel* m, const dprt_half* x_dev, dprt_half* y_dev, int64_t n, cudaStream_t stream,
                 std::string& err);
// One launch for every proxy of a stage: object i owns rows [offsets[i], offsets[i+1]) of the packed query arrays
// (bucket-major, sceneOffset), table[i] names its network. table_dev / offsets_dev are device memory; pairs_upper bounds
// the number of 256-row work units (grid size only). All models of a table use one operand dtype.
MlpGroupEntry mlp_group_entry(const MlpModel* m);
int  mlp_forward_group(const MlpGroupEntry* table_dev, const int32_t* offsets_dev, int S, int64_t pairs_upper, int dtype, const dprt_half* x_dev,
                       dprt_half* y_dev, cudaStream_t stream, std::string& err);
int64_t mlp_macs_per_row(const MlpModel* m);

}  // namespace dprt
