// C-ABI glue of libinf3dp: error reporting, weight packing (FP32 sections), engine dispatch, curvature tail.
#include <mutex>
#include <unordered_map>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace inf3dp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Host mirror of packed headers so that inf3dp_siren_jet never reads the device blob synchronously.  Keyed by the blob's
// device address: an owner that frees or overwrites a blob calls inf3dp_siren_release so that a later blob at the same
// address (e.g. a clone of another network) cannot hit a stale header; the map is also bounded.
static std::mutex g_reg_mu;
static std::unordered_map<const void*, SirenHeader> g_registry;
constexpr size_t kMaxRegistryEntries = 4096;

static void registry_put(const void* packed, const SirenHeader& h) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    if (g_registry.size() >= kMaxRegistryEntries) g_registry.clear();   // entries are re-read from the device on demand
    g_registry[packed] = h;
}

static int lookup_header(const void* packed, SirenHeader* out) {
    {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        auto it = g_registry.find(packed);
        if (it != g_registry.end()) {
            *out = it->second;
            return INF3DP_OK;
        }
    }
    // Blob that was copied or packed by another process image: read its header once (synchronous) and cache.
    SirenHeader h;
    INF3DP_CUDA(cudaMemcpy(&h, packed, sizeof(h), cudaMemcpyDeviceToHost));
    INF3DP_CHECK_ARG(h.magic == kSirenMagic, "packed weights: bad magic (not produced by inf3dp_siren_pack_weights)");
    registry_put(packed, h);
    *out = h;
    return INF3DP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// FP32 packing kernels: fold omega0 (modules.py:34 sin(30 z) == sin(W' a + b') with W' = 30 W, b' = 30 b)
// ---------------------------------------------------------------------------------------------------------
__global__ void pack_first_layer(const float* __restrict__ W, const float* __restrict__ b, int in_features,
                                 int hidden, float scale, float* __restrict__ w0, float* __restrict__ b0) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= hidden) return;
    for (int i = 0; i < 4; ++i) w0[n * 4 + i] = i < in_features ? scale * W[n * in_features + i] : 0.f;
    b0[n] = scale * b[n];
}

__global__ void pack_hidden_layer(const float* __restrict__ W, const float* __restrict__ b, int hidden, float scale,
                                  float* __restrict__ wt, float* __restrict__ bo) {
    // wt[k][n] = scale * W[n][k]
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < hidden * hidden) {
        int k = idx / hidden, n = idx % hidden;
        wt[idx] = scale * W[n * hidden + k];
    }
    if (idx < hidden) bo[idx] = scale * b[idx];
}

__global__ void pack_out_layer(const float* __restrict__ W, const float* __restrict__ b, int hidden, int out,
                               float* __restrict__ wout, float* __restrict__ bout) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < kMaxOut * hidden) {
        int o = idx / hidden;
        wout[idx] = o < out ? W[idx] : 0.f;
    }
    if (idx < kMaxOut) bout[idx] = idx < out ? b[idx] : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// Curvature tail (diff_operators.py:167-220 + :280-298), one thread per point, closed-form symmetric 2x2 eigen.
// ---------------------------------------------------------------------------------------------------------
__global__ void principal_curvatures_kernel(const float* __restrict__ grad, const float* __restrict__ hess,
                                            int64_t n, float* __restrict__ kappa, float* __restrict__ dirs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g[3] = {grad[3 * i], grad[3 * i + 1], grad[3 * i + 2]};
    float Hm[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) Hm[q] = hess[9 * i + q];
    float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    float nx = g[0] / gn, ny = g[1] / gn, nz = g[2] / gn;
    float nn[3] = {nx, ny, nz};
    // argmin |n| with first-index tie break (torch.argmin)
    int mi = 0;
    float mv = fabsf(nx);
    if (fabsf(ny) < mv) { mv = fabsf(ny); mi = 1; }
    if (fabsf(nz) < mv) { mv = fabsf(nz); mi = 2; }
    float a[3] = {0.f, 0.f, 0.f};
    a[mi] = 1.f;
    float dot = nn[mi];
    float e1[3], e2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) e1[d] = a[d] - dot * nn[d];
    float e1n = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
#pragma unroll
    for (int d = 0; d < 3; ++d) e1[d] /= e1n;
    e2[0] = ny * e1[2] - nz * e1[1];
    e2[1] = nz * e1[0] - nx * e1[2];
    e2[2] = nx * e1[1] - ny * e1[0];
    float He1[3], He2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        He1[r] = Hm[3 * r] * e1[0] + Hm[3 * r + 1] * e1[1] + Hm[3 * r + 2] * e1[2];
        He2[r] = Hm[3 * r] * e2[0] + Hm[3 * r + 1] * e2[1] + Hm[3 * r + 2] * e2[2];
    }
    float w11 = -(e1[0] * He1[0] + e1[1] * He1[1] + e1[2] * He1[2]) / gn;
    float w12 = -(e1[0] * He2[0] + e1[1] * He2[1] + e1[2] * He2[2]) / gn;
    float w22 = -(e2[0] * He2[0] + e2[1] * He2[1] + e2[2] * He2[2]) / gn;
    // symmetric 2x2 eigen-decomposition, ascending eigenvalues (torch.linalg.eigh)
    float tr = 0.5f * (w11 + w22), df = 0.5f * (w11 - w22);
    float rad = sqrtf(df * df + w12 * w12);
    float k1 = tr - rad, k2 = tr + rad;
    // eigenvector of k2: (w12, k2 - w11) or (k2 - w22, w12), pick the better conditioned one
    float vx, vy;
    if (df >= 0.f) { vx = k2 - w22; vy = w12; } else { vx = w12; vy = k2 - w11; }
    float vn = sqrtf(vx * vx + vy * vy);
    if (vn > 0.f) { vx /= vn; vy /= vn; } else { vx = 0.f; vy = 1.f; }  // umbilic: any basis (eigh returns identity)
    // v_min is orthogonal to v_max
    float ux = -vy, uy = vx;
    if (vn == 0.f) { ux = 1.f; uy = 0.f; }
    kappa[2 * i] = k1;
    kappa[2 * i + 1] = k2;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        dirs[6 * i + d] = ux * e1[d] + uy * e2[d];
        dirs[6 * i + 3 + d] = vx * e1[d] + vy * e2[d];
    }
}

// Unit normals from gradients: out = g / max(|g|, eps)  (torch.nn.functional.normalize semantics, the step the
// reference applies after field_query_with_grads, e.g. toolpath_utils.py:180-181).  One thread per point.
// the same on packed (value, d/dx, d/dy, d/dz) rows: identical arithmetic, gradient read as one 16-byte load
__global__ void unit_normals_rows_kernel(const float* __restrict__ rows, int64_t n, float eps, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r = __ldg(reinterpret_cast<const float4*>(rows) + i);
    const float x = r.y, y = r.z, z = r.w;
    const float nrm = fmaxf(sqrtf(x * x + y * y + z * z), eps);
    out[3 * i] = x / nrm;
    out[3 * i + 1] = y / nrm;
    out[3 * i + 2] = z / nrm;
}

__global__ void unit_normals_kernel(const float* __restrict__ g, int64_t n, float eps, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = g[3 * i], y = g[3 * i + 1], z = g[3 * i + 2];
    const float nrm = fmaxf(sqrtf(x * x + y * y + z * z), eps);
    out[3 * i] = x / nrm;
    out[3 * i + 1] = y / nrm;
    out[3 * i + 2] = z / nrm;
}

static std::atomic<int> g_siren_sm_limit{0};
int siren_sm_limit() { return g_siren_sm_limit.load(); }

}  // namespace inf3dp

using namespace inf3dp;

extern "C" {

const char* inf3dp_last_error(void) { return g_err; }
int inf3dp_version(void) { return 100; }

int inf3dp_siren_set_sm_limit(int max_sms) {
    INF3DP_CHECK_ARG(max_sms >= 0, "siren_set_sm_limit: negative limit");
    g_siren_sm_limit.store(max_sms);
    return INF3DP_OK;
}

size_t inf3dp_siren_packed_bytes(int in_features, int hidden, int num_hidden_layers, int out_features) {
    if (!siren_config_supported(in_features, hidden, num_hidden_layers, out_features)) return 0;
    return siren_layout(in_features, hidden, num_hidden_layers, out_features, INF3DP_ACT_SINE).total;
}

int inf3dp_siren_pack_weights(const float* const* W, const float* const* b, int in_features, int hidden,
                              int num_hidden_layers, int out_features, int activation, float omega0, void* packed,
                              size_t packed_bytes, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(W && b && packed, "pack_weights: null pointer");
    INF3DP_CHECK_ARG(siren_config_supported(in_features, hidden, num_hidden_layers, out_features),
                     "pack_weights: unsupported configuration in=%d hidden=%d layers=%d out=%d (need hidden=256, "
                     "in in {3,4}, 1<=out<=%d)", in_features, hidden, num_hidden_layers, out_features, kMaxOut);
    INF3DP_CHECK_ARG(activation == INF3DP_ACT_SINE || activation == INF3DP_ACT_RELU, "pack_weights: bad activation");
    const int L = num_hidden_layers;
    SirenLayout lay = siren_layout(in_features, hidden, L, out_features, activation);
    if (packed_bytes < lay.total) {
        set_error("pack_weights: packed buffer has %zu bytes, need %zu", packed_bytes, lay.total);
        return INF3DP_EWORKSPACE;
    }
    for (int i = 0; i < L + 2; ++i) INF3DP_CHECK_ARG(W[i] && b[i], "pack_weights: null layer pointer %d", i);
    char* base = (char*)packed;
    const float scale = activation == INF3DP_ACT_SINE ? omega0 : 1.f;
    cudaStream_t st = (cudaStream_t)stream;

    SirenHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kSirenMagic;
    h.in_features = in_features; h.hidden = hidden; h.num_hidden_layers = L; h.out_features = out_features;
    h.activation = activation; h.omega0 = omega0; h.tc_ok = lay.tc_ok ? 1 : 0;
    h.off_w0 = lay.off_w0; h.off_b0 = lay.off_b0; h.off_wt = lay.off_wt; h.off_b = lay.off_b;
    h.off_wout = lay.off_wout; h.off_bout = lay.off_bout; h.off_tc = lay.tc_ok ? lay.off_tc : 0;
    h.total_bytes = lay.total;
    h.tc_scale = 1.f; h.tc_inv_scale = 1.f;
    h.off_rev = (lay.rev_ok || lay.mlp_ok) ? lay.off_rev : 0; h.rev_ok = 0; h.mlp_ok = 0;

    pack_first_layer<<<(hidden + 255) / 256, 256, 0, st>>>(W[0], b[0], in_features, hidden, scale,
                                                          (float*)(base + lay.off_w0), (float*)(base + lay.off_b0));
    INF3DP_LAUNCH_CHECK("pack_first_layer");
    for (int l = 0; l < L; ++l) {
        pack_hidden_layer<<<(hidden * hidden + 255) / 256, 256, 0, st>>>(
            W[1 + l], b[1 + l], hidden, scale, (float*)(base + lay.off_wt) + (size_t)l * hidden * hidden,
            (float*)(base + lay.off_b) + (size_t)l * hidden);
        INF3DP_LAUNCH_CHECK("pack_hidden_layer");
    }
    pack_out_layer<<<(kMaxOut * hidden + 255) / 256, 256, 0, st>>>(W[L + 1], b[L + 1], hidden, out_features,
                                                                  (float*)(base + lay.off_wout),
                                                                  (float*)(base + lay.off_bout));
    INF3DP_LAUNCH_CHECK("pack_out_layer");
    if (lay.tc_ok) {
        int rc = siren_tc_pack(W, b, omega0, out_features, packed, lay, h, st);
        if (rc != INF3DP_OK) return rc;
        if (lay.rev_ok) {
            rc = siren_rev_pack(W, omega0, siren_tc_maxabs_bits(packed, lay), packed, lay, h, st);
            if (rc != INF3DP_OK) return rc;
        }
    }
    if (lay.mlp_ok) {
        int rc = siren_mlp_pack(W, b, out_features, packed, lay, h, st);
        if (rc != INF3DP_OK) return rc;
    }
    INF3DP_CUDA(cudaMemcpyAsync(packed, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    // the header copy reads a stack object: make sure it has left the host before returning
    INF3DP_CUDA(cudaStreamSynchronize(st));
    registry_put(packed, h);
    return INF3DP_OK;
}

int inf3dp_siren_release(const void* packed) {
    if (packed == nullptr) return INF3DP_OK;
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_registry.erase(packed);
    return INF3DP_OK;
}

size_t inf3dp_siren_jet_workspace_bytes(void) { return siren_rev_workspace_bytes(); }

int inf3dp_siren_jet_ws(const void* packed, const float* x, int64_t n, int order, float* y, float* J, float* H,
                        int mode, void* workspace, size_t workspace_bytes, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(packed != nullptr, "siren_jet: null packed weights");
    INF3DP_CHECK_ARG(n >= 0, "siren_jet: negative point count");
    INF3DP_CHECK_ARG(order >= 0 && order <= 2, "siren_jet: order must be 0, 1 or 2 (got %d)", order);
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(x && y, "siren_jet: null x or y");
    INF3DP_CHECK_ARG(order < 1 || J, "siren_jet: order >= 1 needs J");
    INF3DP_CHECK_ARG(order < 2 || H, "siren_jet: order == 2 needs H");
    SirenHeader h;
    int rc = lookup_header(packed, &h);
    if (rc != INF3DP_OK) return rc;
    INF3DP_CHECK_ARG(order == 0 || h.in_features == 3, "siren_jet: derivatives need in_features == 3");
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc_capable = h.tc_ok && order <= 1;
    const bool rev_capable = h.rev_ok && order == 1;
    const bool jet2_capable = h.rev_ok && order == 2;   // single-output sine network: Hessian on the tensor cores
    const bool mlp_capable = h.mlp_ok && order == 0;    // ReLU level classifier on the tensor cores
    const bool have_ws = workspace != nullptr && workspace_bytes >= siren_rev_workspace_bytes();
    switch (mode) {
        case INF3DP_MODE_AUTO:
            if (rev_capable && have_ws) return siren_rev_launch(packed, h, x, n, y, J, workspace, workspace_bytes, st);
            if (jet2_capable && have_ws) return siren_hess_launch(packed, h, x, n, y, J, H, workspace, workspace_bytes, st);
            if (jet2_capable) return siren_jet2_launch(packed, h, x, n, y, J, H, st);
            if (mlp_capable) return siren_mlp_launch(packed, h, x, n, y, st);
            if (h.rev_ok && order == 0) return siren_value_launch(packed, h, x, n, y, st);   // single-output sine field: CTA-pair engine
            if (tc_capable) return siren_tc_launch(packed, h, x, n, order, y, J, true, st);
            return siren_simt_launch(packed, h, x, n, order, y, J, H, st);
        case INF3DP_MODE_SIMT:
            return siren_simt_launch(packed, h, x, n, order, y, J, H, st);
        case INF3DP_MODE_TC_PRECISE:
            if (jet2_capable) return siren_jet2_launch(packed, h, x, n, y, J, H, st);
            if (mlp_capable) return siren_mlp_launch(packed, h, x, n, y, st);
            /* fallthrough */
        case INF3DP_MODE_TC_FAST:
            INF3DP_CHECK_ARG(tc_capable, "siren_jet: tensor-core engine needs sine, in=3, hidden=256, 3 hidden layers, "
                             "out<=4 and order<=1");
            return siren_tc_launch(packed, h, x, n, order, y, J, mode == INF3DP_MODE_TC_PRECISE, st);
        case INF3DP_MODE_TC_REVERSE:
            if (h.rev_ok && order == 0) return siren_value_launch(packed, h, x, n, y, st);
            if (jet2_capable) return siren_hess_launch(packed, h, x, n, y, J, H, workspace, workspace_bytes, st);
            INF3DP_CHECK_ARG(rev_capable, "siren_jet: the CTA-pair engine needs sine, in=3, hidden=256, 3 hidden layers, "
                             "out=1");
            return siren_rev_launch(packed, h, x, n, y, J, workspace, workspace_bytes, st);
        default:
            set_error("siren_jet: unknown mode %d", mode);
            return INF3DP_EINVAL;
    }
}

int inf3dp_siren_jet(const void* packed, const float* x, int64_t n, int order, float* y, float* J, float* H,
                     int mode, inf3dp_stream_t stream) {
    return inf3dp_siren_jet_ws(packed, x, n, order, y, J, H, mode, nullptr, 0, stream);
}

int inf3dp_siren_value_grad_rows(const void* packed, const float* x, int64_t n, float* rows, void* workspace,
                                 size_t workspace_bytes, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(packed != nullptr, "siren_value_grad_rows: null packed weights");
    INF3DP_CHECK_ARG(n >= 0, "siren_value_grad_rows: negative point count");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(x && rows, "siren_value_grad_rows: null x or rows");
    SirenHeader h;
    int rc = lookup_header(packed, &h);
    if (rc != INF3DP_OK) return rc;
    INF3DP_CHECK_ARG(h.rev_ok, "siren_value_grad_rows: needs a single-output sine network 3-256-256-256-256-1");
    return siren_rev_launch_rows(packed, h, x, n, rows, workspace, workspace_bytes, (cudaStream_t)stream);
}

int inf3dp_siren_value_grad_rows_peers(const void* packed, const float* x, int64_t n, float* const* peer_rows, int n_peers,
                                       float* multicast_rows, void* workspace, size_t workspace_bytes,
                                       inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(packed != nullptr, "siren_value_grad_rows_peers: null packed weights");
    INF3DP_CHECK_ARG(n >= 0, "siren_value_grad_rows_peers: negative point count");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(x != nullptr, "siren_value_grad_rows_peers: null x");
    SirenHeader h;
    int rc = lookup_header(packed, &h);
    if (rc != INF3DP_OK) return rc;
    INF3DP_CHECK_ARG(h.rev_ok, "siren_value_grad_rows_peers: needs a single-output sine network 3-256-256-256-256-1");
    return siren_rev_launch_rows_peers(packed, h, x, n, peer_rows, n_peers, multicast_rows, workspace, workspace_bytes,
                                       (cudaStream_t)stream);
}

int inf3dp_unit_normals_rows(const float* rows, int64_t n, float eps, float* normals, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0, "unit_normals_rows: negative n");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(rows && normals, "unit_normals_rows: null pointer");
    INF3DP_CHECK_ARG((reinterpret_cast<uintptr_t>(rows) & 15) == 0, "unit_normals_rows: rows must be 16-byte aligned");
    unit_normals_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, n, eps, normals);
    INF3DP_LAUNCH_CHECK("unit_normals_rows_kernel");
    return INF3DP_OK;
}

int inf3dp_unit_normals(const float* grad, int64_t n, float eps, float* normals, inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0, "unit_normals: negative n");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(grad && normals, "unit_normals: null pointer");
    unit_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad, n, eps, normals);
    INF3DP_LAUNCH_CHECK("unit_normals_kernel");
    return INF3DP_OK;
}

int inf3dp_principal_curvatures(const float* grad, const float* hess, int64_t n, float* kappa, float* dirs,
                                inf3dp_stream_t stream) {
    INF3DP_CHECK_ARG(n >= 0, "principal_curvatures: negative n");
    if (n == 0) return INF3DP_OK;
    INF3DP_CHECK_ARG(grad && hess && kappa && dirs, "principal_curvatures: null pointer");
    int64_t blocks = (n + 255) / 256;
    principal_curvatures_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad, hess, n, kappa, dirs);
    INF3DP_LAUNCH_CHECK("principal_curvatures_kernel");
    return INF3DP_OK;
}

}  // extern "C"
