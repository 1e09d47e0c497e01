// dist.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference has no distributed path; its DSGD mode runs g threads over g x g blocks
// (BiasedMatrixFactorization.cs:205-215). Lifted to GPUs: rank r owns user block r for good and holds
// item block (S + r) mod R during GPU-level sub-epoch S; after it the item block (factors + biases)
// moves to rank r - 1 with one grouped ncclSend / ncclRecv (see sgd.cu).
#include "common.cuh"
#include <nccl.h>
#include <new>

using namespace mml;

#define MML_NCCL(expr)                                                                          \
    do {                                                                                        \
        ncclResult_t _r = (expr);                                                               \
        if (_r != ncclSuccess) {                                                                \
            mml::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r)); \
            return MML_ERR_NCCL;                                                                \
        }                                                                                       \
    } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "mml_dist_unique_id hands out 128 bytes");

extern "C" int32_t mml_dist_unique_id(uint8_t* out128)
{
    MML_CHECK(out128 != nullptr, MML_ERR_ARG, "mml_dist_unique_id: NULL argument");
    ncclUniqueId id;
    MML_NCCL(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return MML_OK;
}

namespace mml { int32_t ctx_create_on_device(int dev, mml_ctx** out); }

extern "C" int32_t mml_ctx_create_dist(int32_t rank, int32_t world, int32_t device, const uint8_t* unique_id128, mml_ctx** out)
{
    MML_CHECK(out && unique_id128, MML_ERR_ARG, "mml_ctx_create_dist: NULL argument");
    MML_CHECK(world >= 1 && rank >= 0 && rank < world, MML_ERR_ARG, "mml_ctx_create_dist: rank %d of %d", rank, world);
    MML_TRY(ctx_create_on_device(device, out));
    Ctx* c = ctx_of(*out);
    c->rank = rank; c->n_gpus = world;
    if (world > 1) {
        ncclUniqueId id;
        memcpy(&id, unique_id128, sizeof(id));
        ncclComm_t comm;
        ncclResult_t r = ncclCommInitRank(&comm, world, id, rank);
        if (r != ncclSuccess) {
            set_error("ncclCommInitRank failed: %s", ncclGetErrorString(r));
            mml_ctx_destroy(*out); *out = nullptr;
            return MML_ERR_NCCL;
        }
        c->comm = (void*)comm;
    }
    return MML_OK;
}

namespace mml {

// One process, several GPUs: all communicators at once (ncclCommInitAll), rank r = position in `peers`.
int32_t dist_init_all(std::vector<Ctx*>& peers)
{
    const int n = (int)peers.size();
    std::vector<int> devs((size_t)n);
    for (int r = 0; r < n; r++) devs[(size_t)r] = peers[(size_t)r]->device;
    std::vector<ncclComm_t> comms((size_t)n);
    MML_NCCL(ncclCommInitAll(comms.data(), n, devs.data()));
    for (int r = 0; r < n; r++) { peers[(size_t)r]->comm = (void*)comms[(size_t)r]; peers[(size_t)r]->rank = r; peers[(size_t)r]->n_gpus = n; }
    return MML_OK;
}

int32_t dist_destroy(Ctx* c)
{
    if (c->comm) { ncclCommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
    return MML_OK;
}

// in-place sum over ranks
int32_t dist_allreduce_u32(Ctx* c, uint32_t* d_buf, size_t n)
{
    if (c->n_gpus <= 1) return MML_OK;
    MML_NCCL(ncclAllReduce(d_buf, d_buf, n, ncclUint32, ncclSum, (ncclComm_t)c->comm, c->stream));
    return MML_OK;
}

int32_t dist_allreduce_f64(Ctx* c, double* d_buf, size_t n)
{
    if (c->n_gpus <= 1) return MML_OK;
    MML_NCCL(ncclAllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
    return MML_OK;
}

int32_t dist_allreduce_f64_max(Ctx* c, double* d_buf, size_t n)
{
    if (c->n_gpus <= 1) return MML_OK;
    MML_NCCL(ncclAllReduce(d_buf, d_buf, n, ncclDouble, ncclMax, (ncclComm_t)c->comm, c->stream));
    return MML_OK;
}

// One ring step: send [send_a, send_b] to `to`, receive into [recv_a, recv_b] from `from`, grouped.
int32_t dist_ring_exchange(Ctx* c, cudaStream_t stream, const float* send_a, size_t n_send_a, const float* send_b, size_t n_send_b, int to,
                           float* recv_a, size_t n_recv_a, float* recv_b, size_t n_recv_b, int from)
{
    if (c->n_gpus <= 1) return MML_OK;
    ncclComm_t comm = (ncclComm_t)c->comm;
    MML_NCCL(ncclGroupStart());
    if (n_send_a) MML_NCCL(ncclSend(send_a, n_send_a, ncclFloat, to, comm, stream));
    if (n_send_b) MML_NCCL(ncclSend(send_b, n_send_b, ncclFloat, to, comm, stream));
    if (n_recv_a) MML_NCCL(ncclRecv(recv_a, n_recv_a, ncclFloat, from, comm, stream));
    if (n_recv_b) MML_NCCL(ncclRecv(recv_b, n_recv_b, ncclFloat, from, comm, stream));
    MML_NCCL(ncclGroupEnd());
    return MML_OK;
}

int32_t dist_broadcast_f32(Ctx* c, float* d_buf, size_t n, int root)
{
    if (c->n_gpus <= 1 || n == 0) return MML_OK;
    MML_NCCL(ncclBroadcast(d_buf, d_buf, n, ncclFloat, root, (ncclComm_t)c->comm, c->stream));
    return MML_OK;
}

int32_t dist_group_start() { MML_NCCL(ncclGroupStart()); return MML_OK; }
int32_t dist_group_end() { MML_NCCL(ncclGroupEnd()); return MML_OK; }

}  // namespace mml
