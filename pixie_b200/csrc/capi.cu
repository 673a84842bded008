// extern "C" boundary of libpixie_b200.so (include/pixie_b200.h). No torch types cross it.
#include "../../include/pixie_b200.h"
#include "mpm.cuh"
#include "unet.cuh"
#include "unet_kernels.cuh"
#include "field_transfer.cuh"

#include <string>

namespace {
thread_local std::string g_err;
int set_err(const std::string& e) { g_err = e; return 1; }
}  // namespace

struct pixie_unet_s { pixie::UNet* u; };
struct pixie_mpm_s { pixie::Mpm* m; };

extern "C" {

const char* pixie_last_error(void) { return g_err.c_str(); }
int pixie_abi_version(void) { return 1; }

int pixie_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return major == 10 ? 1 : 0;
}

static int require_device() {
    if (!pixie_device_ok()) return set_err("pixie_b200 requires an sm_100 (B200) CUDA device; there is no CPU fallback");
    return 0;
}

// ------------------------------------------------------------------------------------- U-Net
int pixie_unet_create(const pixie_unet_config* cfg, pixie_unet_t* out) {
    if (!cfg || !out) return set_err("null argument");
    if (require_device()) return 1;
    std::string e;
    pixie::UNet* u = pixie::unet_create(*cfg, e);
    if (!u) return set_err(e);
    *out = new pixie_unet_s{u};
    return 0;
}
int pixie_unet_set_tensor(pixie_unet_t h, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    if (!h || !name || !host_data) return set_err("null argument");
    if (pixie::unet_set_tensor(h->u, name, host_data, shape, ndim)) return set_err(pixie::unet_error(h->u));
    return 0;
}
int pixie_unet_finalize(pixie_unet_t h) {
    if (!h) return set_err("null handle");
    if (pixie::unet_finalize(h->u)) return set_err(pixie::unet_error(h->u));
    return 0;
}
int pixie_unet_forward(pixie_unet_t h, const void* feat, int batch, float* out, void* stream) {
    if (!h || !feat || !out) return set_err("null argument");
    if (pixie::unet_forward(h->u, feat, batch, out, (cudaStream_t)stream)) return set_err(pixie::unet_error(h->u));
    return 0;
}
int pixie_unet_forward_ncdhw(pixie_unet_t h, const float* feat, int batch, float* out, void* stream) {
    if (!h || !feat || !out) return set_err("null argument");
    if (pixie::unet_forward_ncdhw(h->u, feat, batch, out, (cudaStream_t)stream)) return set_err(pixie::unet_error(h->u));
    return 0;
}
int pixie_unet_forward_host(pixie_unet_t h, const void* feat, int batch, float* out, void* stream) {
    if (!h || !feat || !out) return set_err("null argument");
    if (pixie::unet_forward_host(h->u, feat, batch, out, (cudaStream_t)stream)) return set_err(pixie::unet_error(h->u));
    return 0;
}
int pixie_unet_profile(pixie_unet_t h, const void* feat, int batch, float* out, void* stream, float* ms, int* kinds, double* flops, int cap) {
    if (!h || !feat || !out || !ms || !kinds || !flops) { set_err("null argument"); return -1; }
    const int n = pixie::unet_profile(h->u, feat, batch, out, (cudaStream_t)stream, ms, kinds, flops, cap);
    if (n < 0) set_err(pixie::unet_error(h->u));
    return n;
}
int pixie_field_extract(const float* pred, int n_classes, const float* mask, int D, const double ranges[6], const double bmin[3], const double bmax[3],
                        float* pos, float* density, float* E, float* nu, int* material, float* conf, int* count_host, void* stream) {
    if (!pred || !mask || !ranges || !bmin || !bmax || !pos || !density || !E || !nu || !material || !conf || !count_host) return set_err("null argument");
    if (D < 2 || n_classes < 1) return set_err("field_extract: need D >= 2 and at least one class channel");
    if (require_device()) return 1;
    if (pixie::field_extract(pred, n_classes, mask, D, ranges, bmin, bmax, pos, density, E, nu, material, conf, count_host, (cudaStream_t)stream))
        return set_err("field_extract failed");
    return 0;
}
int pixie_knn_assign(const float* query, int nq, const float* pos, const float* density, const float* E, const float* nu, const int* material,
                     const int* part, const float* conf, int m, int k, float threshold, int weighted, const float defaults[4], int def_material,
                     int def_part, float* o_density, float* o_E, float* o_nu, int* o_material, int* o_part, float* o_conf, int* n_too_far_host,
                     void* stream) {
    if (!query || !pos || !density || !E || !nu || !material || !part || !conf || !defaults || !o_density || !o_E || !o_nu || !o_material ||
        !o_part || !o_conf || !n_too_far_host) return set_err("null argument");
    if (k < 1 || k > 16) return set_err("knn_assign: k must be in [1, 16]");
    if (m < 1) return set_err("knn_assign: empty material point cloud");
    if (require_device()) return 1;
    if (pixie::knn_assign(query, nq, pos, density, E, nu, material, part, conf, m, k, threshold, weighted, defaults, def_material, def_part,
                          o_density, o_E, o_nu, o_material, o_part, o_conf, n_too_far_host, (cudaStream_t)stream))
        return set_err("knn_assign failed");
    return 0;
}
int pixie_particle_volume(const float* pos, int n, int grid_n, float grid_dx, float* vol, void* stream) {
    if (!pos || !vol) return set_err("null argument");
    if (grid_n < 1 || !(grid_dx > 0.f)) return set_err("particle_volume: bad grid");
    if (require_device()) return 1;
    if (pixie::particle_volume(pos, n, grid_n, grid_dx, vol, (cudaStream_t)stream)) return set_err("particle_volume failed");
    return 0;
}
int pixie_frame_transform(const float* pos, const float* cov, int n, float z_shift, float scale, const float mean[3], const float* rotations,
                          int n_rot, float* pos_out, float* cov_out, void* stream) {
    if (!pos || !mean || !pos_out || (cov && !cov_out) || (n_rot > 0 && !rotations)) return set_err("null argument");
    if (n_rot < 0 || n_rot > 8) return set_err("frame_transform: at most 8 rotations");
    if (require_device()) return 1;
    if (pixie::frame_transform(pos, cov, n, z_shift, scale, mean, rotations, n_rot, pos_out, cov_out, (cudaStream_t)stream)) return set_err("frame_transform failed");
    return 0;
}
int pixie_pack_predictions(const float* seg_logits_dev, const float* cont_dev, float* out_dev, int batch, int64_t voxels, int n_classes, void* stream) {
    if (!seg_logits_dev || !cont_dev || !out_dev) return set_err("null argument");
    if (require_device()) return 1;
    if (pixie::launch_pack_predictions(seg_logits_dev, cont_dev, out_dev, batch, voxels, n_classes, (cudaStream_t)stream)) return set_err("launch failed");
    return 0;
}
int pixie_unet_launch_count(pixie_unet_t h) { return h ? pixie::unet_launch_count(h->u) : 0; }
double pixie_unet_flops(pixie_unet_t h) { return h ? pixie::unet_flops(h->u) : 0.0; }
int pixie_unet_check(pixie_unet_t h) {
    if (!h) return set_err("null handle");
    if (pixie::unet_check(h->u)) return set_err(pixie::unet_error(h->u));
    return 0;
}
int64_t pixie_unet_debug_fetch(pixie_unet_t h, const char* name, float* host_out, int64_t capacity) {
    if (!h || !name) { set_err("null argument"); return -1; }
    const int64_t n = pixie::unet_debug_fetch(h->u, name, host_out, capacity);
    if (n < 0) set_err(pixie::unet_error(h->u));
    return n;
}
void pixie_unet_destroy(pixie_unet_t h) {
    if (!h) return;
    pixie::unet_destroy(h->u);
    delete h;
}

// ------------------------------------------------------------------------------------- MPM
int pixie_mpm_create(int n_particles, int n_grid, float grid_lim, pixie_mpm_t* out) {
    if (!out) return set_err("null argument");
    if (require_device()) return 1;
    std::string e;
    pixie::Mpm* m = pixie::mpm_create(n_particles, n_grid, grid_lim, e);
    if (!m) return set_err(e);
    *out = new pixie_mpm_s{m};
    return 0;
}
#define MPM_CALL(expr) do { if (!h) return set_err("null handle"); if (expr) return set_err(pixie::mpm_error(h->m)); return 0; } while (0)
int pixie_mpm_bind(pixie_mpm_t h, int field, void* dev_ptr) { MPM_CALL(pixie::mpm_bind(h->m, field, dev_ptr)); }
int pixie_mpm_set_params(pixie_mpm_t h, const pixie_mpm_params* p) { if (!p) return set_err("null params"); MPM_CALL(pixie::mpm_set_params(h->m, *p)); }
int pixie_mpm_add_bc(pixie_mpm_t h, const pixie_mpm_bc* bc) { if (!bc) return set_err("null bc"); MPM_CALL(pixie::mpm_add_bc(h->m, *bc)); }
int pixie_mpm_clear_bcs(pixie_mpm_t h) { MPM_CALL(pixie::mpm_clear_bcs(h->m)); }
int pixie_mpm_set_time(pixie_mpm_t h, double t) { MPM_CALL(pixie::mpm_set_time(h->m, t)); }
int pixie_mpm_get_time(pixie_mpm_t h, double* t) { MPM_CALL(pixie::mpm_get_time(h->m, t)); }
int pixie_mpm_step(pixie_mpm_t h, int n, double dt, void* stream) { MPM_CALL(pixie::mpm_step(h->m, n, dt, (cudaStream_t)stream)); }
int pixie_mpm_compute_mu_lam(pixie_mpm_t h, void* s) { MPM_CALL(pixie::mpm_compute_mu_lam(h->m, (cudaStream_t)s)); }
int pixie_mpm_compute_bulk(pixie_mpm_t h, void* s) { MPM_CALL(pixie::mpm_compute_bulk(h->m, (cudaStream_t)s)); }
int pixie_mpm_compute_mass(pixie_mpm_t h, void* s) { MPM_CALL(pixie::mpm_compute_mass(h->m, (cudaStream_t)s)); }
int pixie_mpm_compute_cov_from_F(pixie_mpm_t h, void* s) { MPM_CALL(pixie::mpm_compute_cov_from_F(h->m, (cudaStream_t)s)); }
int pixie_mpm_compute_R_from_F(pixie_mpm_t h, void* s) { MPM_CALL(pixie::mpm_compute_R_from_F(h->m, (cudaStream_t)s)); }
int pixie_mpm_apply_additional_params(pixie_mpm_t h, const float* boxes, int n_boxes, void* s) {
    MPM_CALL(pixie::mpm_apply_additional_params(h->m, boxes, n_boxes, (cudaStream_t)s));
}
int pixie_mpm_select_box(pixie_mpm_t h, const float point[3], const float size[3], int* mask, void* s) {
    MPM_CALL(pixie::mpm_select_box(h->m, point, size, mask, (cudaStream_t)s));
}
int pixie_mpm_select_cylinder(pixie_mpm_t h, const float point[3], const float normal[3], float hh, float radius, int* mask, void* s) {
    MPM_CALL(pixie::mpm_select_cylinder(h->m, point, normal, hh, radius, mask, (cudaStream_t)s));
}
int pixie_mpm_sync(pixie_mpm_t h, void* s) { MPM_CALL(pixie::mpm_sync(h->m, (cudaStream_t)s)); }
int pixie_mpm_set_active_count(pixie_mpm_t h, int n_active) { MPM_CALL(pixie::mpm_set_active_count(h->m, n_active)); }
int pixie_mpm_grid_ptrs(pixie_mpm_t h, float** mv4, float** v4) { MPM_CALL(pixie::mpm_grid_ptrs(h->m, mv4, v4)); }
int pixie_mpm_exchange_buffer(pixie_mpm_t h, void** base, size_t* bytes) { MPM_CALL(pixie::mpm_exchange_buffer(h->m, base, bytes)); }
int pixie_mpm_slab_attach(pixie_mpm_t h, int x0, int x1, int slack, const void* left, const void* right) {
    MPM_CALL(pixie::mpm_slab_attach(h->m, x0, x1, slack, left, right));
}
int pixie_mpm_slab_phase(pixie_mpm_t h, int phase, double dt, void* s) { MPM_CALL(pixie::mpm_slab_phase(h->m, phase, dt, (cudaStream_t)s)); }
int pixie_mpm_slab_error(pixie_mpm_t h, int* flag) { MPM_CALL(pixie::mpm_slab_error(h->m, flag)); }
int pixie_mpm_slab_excursion(pixie_mpm_t h, int* d_out, void* s) { MPM_CALL(pixie::mpm_slab_excursion(h->m, d_out, (cudaStream_t)s)); }
int pixie_ipc_export(const void* dev_ptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, const_cast<void*>(dev_ptr)) != cudaSuccess) return set_err(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError()));
    memcpy(handle, &hd, 64);
    return 0;
}
int pixie_ipc_open(const unsigned char handle[64], void** dev_ptr) {
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    if (cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        return set_err(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(cudaGetLastError()));
    return 0;
}
int pixie_ipc_close(void* dev_ptr) { return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? 0 : set_err("cudaIpcCloseMemHandle failed"); }
long long pixie_mpm_launch_count(pixie_mpm_t h) { return h ? pixie::mpm_launch_count(h->m) : 0; }
void pixie_mpm_destroy(pixie_mpm_t h) {
    if (!h) return;
    pixie::mpm_destroy(h->m);
    delete h;
}

}  // extern "C"
