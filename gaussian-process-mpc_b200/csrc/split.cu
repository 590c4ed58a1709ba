// One rollout split over several GPUs: connection management (include/gpmpc.h, gpmpc_split_*).
//
// The horizon is sequential and the reference evaluates one control sequence per IPOPT callback (src/mpc.py:202-255),
// so the only way several GPUs can shorten ONE evaluation is to split the pair space of every step.  Each
// rank keeps a replica of the fitted GP, sweeps 1/world of the tile list in the persistent kernel
// (mm_rollout_single.cuh) and, once per step, writes its E (1 + 2D) x 2 sums straight into every peer's mailbox with
// NVLink P2P stores followed by a release flag; no host round trip, no NCCL call per step.  This file only creates the
// mailbox and maps the peers' mailboxes: across processes through CUDA IPC handles (one process per GPU, the handles
// travel through torch.distributed), or directly for handles that live in the same process.
#include "common.cuh"

using namespace gpmpc;

static size_t mailbox_bytes()
{
    return (size_t)kSplitMaxWorld * 2 * kSplitNV * sizeof(double) + (size_t)kSplitMaxWorld * 2 * sizeof(unsigned long long);
}
static unsigned long long *flags_of(double *mail) { return reinterpret_cast<unsigned long long *>(mail + (size_t)kSplitMaxWorld * 2 * kSplitNV); }

static int ensure_mailbox(gpmpc_ctx *h)
{
    if (h->split_buf.p) return GPMPC_OK;
    GP_CUDA(h, cudaSetDevice(h->device));
    GP_CUDA(h, h->split_buf.reserve(mailbox_bytes()));
    GP_CUDA(h, cudaMemset(h->split_buf.p, 0, mailbox_bytes()));
    return GPMPC_OK;
}

extern "C" int gpmpc_split_export(gpmpc_handle h, void *ipc_handle_out)
{
    if (!h || !ipc_handle_out) return GPMPC_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == GPMPC_IPC_HANDLE_BYTES, "IPC handle size");
    int rc = ensure_mailbox(h);
    if (rc) return rc;
    cudaIpcMemHandle_t ih;
    GP_CUDA(h, cudaIpcGetMemHandle(&ih, h->split_buf.p));
    std::memcpy(ipc_handle_out, &ih, sizeof ih);
    return GPMPC_OK;
}

extern "C" int gpmpc_split_disconnect(gpmpc_handle h)
{
    if (!h) return GPMPC_ERR_INVALID;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int r = 0; r < kSplitMaxWorld; ++r) {
        if (h->split_ipc[r] && h->peer_mail[r]) cudaIpcCloseMemHandle(h->peer_mail[r]);
        h->peer_mail[r] = nullptr; h->peer_flags[r] = nullptr; h->split_ipc[r] = false;
    }
    h->split_world = 1; h->split_rank = 0;
    return GPMPC_OK;
}

static int split_begin(gpmpc_ctx *h, int rank, int world)
{
    if (world < 1 || world > kSplitMaxWorld || rank < 0 || rank >= world)
        return fail(h, GPMPC_ERR_INVALID, "gpmpc_split_connect: need 1 <= world <= 8 and 0 <= rank < world");
    gpmpc_split_disconnect(h);
    int rc = ensure_mailbox(h);
    if (rc) return rc;
    // a fresh connection starts a fresh sequence: clear the flags (the peers do the same before anybody launches)
    GP_CUDA(h, cudaMemset(h->split_buf.p, 0, mailbox_bytes()));
    h->split_seq = 0;
    h->peer_mail[rank] = h->split_buf.as<double>();
    h->peer_flags[rank] = flags_of(h->peer_mail[rank]);
    return GPMPC_OK;
}

extern "C" int gpmpc_split_connect(gpmpc_handle h, int rank, int world, const void *all_handles)
{
    if (!h || !all_handles) return GPMPC_ERR_INVALID;
    int rc = split_begin(h, rank, world);
    if (rc) return rc;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        cudaIpcMemHandle_t ih;
        std::memcpy(&ih, static_cast<const char *>(all_handles) + (size_t)r * sizeof ih, sizeof ih);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            gpmpc_split_disconnect(h);
            return fail(h, GPMPC_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        }
        h->peer_mail[r] = static_cast<double *>(p);
        h->peer_flags[r] = flags_of(h->peer_mail[r]);
        h->split_ipc[r] = true;
    }
    h->split_world = world; h->split_rank = rank;
    return GPMPC_OK;
}

extern "C" int gpmpc_split_connect_local(gpmpc_handle h, int rank, int world, gpmpc_handle *peers)
{
    if (!h || !peers) return GPMPC_ERR_INVALID;
    int rc = split_begin(h, rank, world);
    if (rc) return rc;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        gpmpc_ctx *p = peers[r];
        if (!p || p->device == h->device) { gpmpc_split_disconnect(h); return fail(h, GPMPC_ERR_INVALID, "gpmpc_split_connect_local: peers must be handles on other devices"); }
        if ((rc = ensure_mailbox(p))) { gpmpc_split_disconnect(h); return rc; }
        GP_CUDA(h, cudaSetDevice(h->device));
        int can = 0;
        cudaDeviceCanAccessPeer(&can, h->device, p->device);
        if (!can) { gpmpc_split_disconnect(h); return fail(h, GPMPC_ERR_UNSUPPORTED, "no peer access between the devices"); }
        cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { gpmpc_split_disconnect(h); return fail(h, GPMPC_ERR_CUDA, cudaGetErrorString(e)); }
        cudaGetLastError();
        h->peer_mail[r] = p->split_buf.as<double>();
        h->peer_flags[r] = flags_of(h->peer_mail[r]);
    }
    h->split_world = world; h->split_rank = rank;
    return GPMPC_OK;
}
