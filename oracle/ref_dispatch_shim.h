/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 * Forced-include (-include) shim used by oracle/build_ref.py when it compiles the reference's OWN,
 * unmodified sources (app/utils/base/cuda/{render_utils,total_variation}{.cpp,_kernel.cu}) from where
 * they lie under /root/reference.  The reference was written against an older PyTorch whose
 * AT_DISPATCH_FLOATING_TYPES accepted `tensor.type()` (an at::DeprecatedTypeProperties); torch 2.11's
 * macro wants a c10::ScalarType.  This shim restores the old calling convention and changes nothing else:
 * every kernel body, launch shape and host function is the reference's.
 */
#pragma once
#include <torch/extension.h>

namespace esr_oracle_shim {
inline c10::ScalarType to_scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
inline c10::ScalarType to_scalar_type(c10::ScalarType t) { return t; }
}  // namespace esr_oracle_shim

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(esr_oracle_shim::to_scalar_type(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
