// C-ABI of libfermi_pt_b200.so (declared in include/fermi_pt_b200.h) and the host-side driver of the (T) path:
// device buffer pool, sharded host->device staging (pinned bounce ring, one PCIe link per GPU), NCCL all-gather of the raw
// arrays over NVLink, layout prep / DF assembly launches, the persistent fused kernel, the scalar all-reduce of E(T).
// Replaces the driver loop + accumulation of
//   src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:20-150 (reference, Julia threads)
// No CPU fallback: every entry point fails loudly without a CUDA device.
#include <cuda_runtime.h>
#include <sched.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fpt_internal.h"
#include "fpt_aux_kernels.cuh"
#include "fpt_gemm.cuh"
#include "fpt_triples.cuh"
#ifdef FPT_WITH_VARIANT2
#include "fpt_triples2.cuh"
#endif

using namespace fpt;
typedef std::chrono::steady_clock wall;
static double ms_since(wall::time_point t0) { return std::chrono::duration<double, std::milli>(wall::now() - t0).count(); }

extern "C" const char* fpt_last_error(void) { return err_text().c_str(); }
extern "C" const char* fpt_version(void) { return "fermi_pt_b200 0.5 (sm_100a)"; }

// The driver is one translation unit (the kernels are templates in headers); its parts, in order:
#include "fpt_api_handle.inl"
#include "fpt_api_staging.inl"
#include "fpt_api_upload.inl"
#include "fpt_api_compute.inl"
#include "fpt_api_ring.inl"
#include "fpt_api_extras.inl"
#include "fpt_api_ao.inl"
#include "fpt_api_diag.inl"
