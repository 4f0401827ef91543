// trx_rollout.h -- internal interface between the objective glue (trx_engine.cu) and the re-rolled
// rollout objective (trx_rollout.cu).  Not part of the C ABI (include/trx_b200.h is).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/trx_b200.h"

struct trx_rollout;

trx_status trxRolloutCreate(const char *options, uint64_t lanes, uint64_t max_device_bytes, int device,
                            bool scalar_rows, int checkpoint_mode, trx_rollout **out, std::string &err);
void trxRolloutDestroy(trx_rollout *ro);
// ~0 for an unknown key
uint64_t trxRolloutInfo(const trx_rollout *ro, const char *key);
trx_status trxRolloutParameterize(trx_rollout *ro, const void *packed, uint64_t bytes, cudaStream_t stream,
                                  std::string &err);
trx_status trxRolloutEvaluate(trx_rollout *ro, const double *d_x, double *d_out, double *d_r, const double *d_seed,
                              bool want_grad, cudaStream_t stream, std::string &err);
uint64_t trxRolloutLaunches(const trx_rollout *ro);
