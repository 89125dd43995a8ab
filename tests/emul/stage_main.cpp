// CPU harness for the host-side staging logic (fermi.jl_b200/csrc/fpt_stage.h): the views the uploads build and the packed-stream copy
// the staging threads run, without a GPU.  Used by tests/test_stage_views.py.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../fermi.jl_b200/csrc/fpt_stage.h"

using namespace fpt;

// kind 0: OVVV chunk (p0, np, c0, cn, half = flag); 1: T2 half; 2: OVOV half.  Copies bytes [off, off + nb) of the packed stream into out
// (piece by piece of `piece` bytes, as the staging threads do) and returns the stream's total size in bytes (out may be NULL).
extern "C" long long fpt_test_view_copy(int kind, int o, int v, int p0, int np, int c0, int cn, int flag, const double* src, long long off,
                                        long long nb, long long piece, int nt, double* out)
{
    const View vw = kind == 0 ? view_ovvv_chunk(o, v, p0, np, c0, cn, flag != 0) : (kind == 1 ? view_t2_half(o, v) : view_ovov_half(o, v));
    if (!out) return (long long)vw.total;
    void* slot = nullptr;
    if (posix_memalign(&slot, 64, (size_t)piece)) return -1;
    for (long long done = 0; done < nb; done += piece) {
        const long long n = nb - done < piece ? nb - done : piece;
        copy_view_to_pinned((char*)slot, (const char*)src, vw, (size_t)(off + done), (size_t)n, nt != 0);
        memcpy((char*)out + done, slot, (size_t)n);
    }
    free(slot);
    return (long long)vw.total;
}
