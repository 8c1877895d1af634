// Temporary stubs for operations whose kernels are not written yet (replaced file by file).
#include "../../include/linalg_b200.h"
#include "ops.cuh"
namespace lq {
#define STUB(name) set_error(c, name ": not implemented yet"); return LQ_ERR_UNSUPPORTED
int large_mgs_qr(Ctx* c, const double*, int, int, int, double*, double*, int*) { STUB("large_mgs_qr"); }
int large_lstsq_mgs(Ctx* c, const double*, const double*, int, int, int, double*, int*) { STUB("large_lstsq_mgs"); }
bool lstsq_stream_kernel_supported(int, int, int) { return false; }
int lstsq_stream_kernel_launch(Ctx*, cudaStream_t, const double*, const double*, long long, int, int, int, double*) { return LQ_ERR_UNSUPPORTED; }
int gram(Ctx* c, const double*, long long, int, double*) { STUB("gram"); }
int eigh_jacobi(Ctx* c, const double*, int, double*, double*) { STUB("eigh"); }
int svd_gram_local(Ctx* c, const double*, long long, int, double, double*, double*, double*, int*, bool) { STUB("svd"); }
int tsqr_local(Ctx* c, const double*, long long, int, double*, double*, bool) { STUB("tsqr"); }
int comm_allreduce_sum(Ctx* c, double*, long long) { STUB("comm"); }
int comm_allgather(Ctx* c, const double*, double*, long long) { STUB("comm"); }
}
using namespace lq;
extern "C" {
int lq_mgs_qr_dev(lq_ctx* h, const double*, int, int, int, double*, double*, int32_t*) { Ctx* c = as_ctx(h); STUB("mgs"); }
int lq_mgs_qr(lq_ctx* h, const double*, int, int, int, double*, double*, int32_t*) { Ctx* c = as_ctx(h); STUB("mgs"); }
int lq_svd_gram_dev(lq_ctx* h, const double*, int64_t, int, double, double*, double*, double*, int*) { Ctx* c = as_ctx(h); STUB("svd"); }
int lq_svd_gram(lq_ctx* h, const double*, int64_t, int, double, double*, double*, double*, int*) { Ctx* c = as_ctx(h); STUB("svd"); }
int lq_gram_dev(lq_ctx* h, const double*, int64_t, int, double*) { Ctx* c = as_ctx(h); STUB("gram"); }
int lq_eigh_dev(lq_ctx* h, const double*, int, double*, double*) { Ctx* c = as_ctx(h); STUB("eigh"); }
int lq_tsqr_dev(lq_ctx* h, const double*, int64_t, int, double*, double*) { Ctx* c = as_ctx(h); STUB("tsqr"); }
int lq_tsqr(lq_ctx* h, const double*, int64_t, int, double*, double*) { Ctx* c = as_ctx(h); STUB("tsqr"); }
int lq_comm_unique_id(void*) { return LQ_ERR_UNSUPPORTED; }
int lq_comm_init(lq_ctx* h, int, int, const void*) { Ctx* c = as_ctx(h); STUB("comm"); }
int lq_comm_destroy(lq_ctx*) { return LQ_OK; }
int lq_comm_allreduce_sum(lq_ctx* h, double*, int64_t) { Ctx* c = as_ctx(h); STUB("comm"); }
int lq_comm_allgather(lq_ctx* h, const double*, double*, int64_t) { Ctx* c = as_ctx(h); STUB("comm"); }
int lq_tsqr_sharded_dev(lq_ctx* h, const double*, int64_t, int, double*, double*) { Ctx* c = as_ctx(h); STUB("tsqr"); }
int lq_svd_gram_sharded_dev(lq_ctx* h, const double*, int64_t, int, double, double*, double*, double*, int*) { Ctx* c = as_ctx(h); STUB("svd"); }
}
