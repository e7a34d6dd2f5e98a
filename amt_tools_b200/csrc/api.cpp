// extern "C" surface of libamtfeat.so (declared in include/amtfeat.h).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <algorithm>

#include "device_guard.h"
#include "plan.h"

struct amtfeat_plan {
    amtfeat::Plan p;
};

namespace amtfeat {
const char *last_error_cstr();
}

using amtfeat::Plan;

extern "C" {

int amtfeat_version(void) { return AMTFEAT_VERSION; }
const char *amtfeat_last_error(void) { return amtfeat::last_error_cstr(); }

int amtfeat_plan_create(const amtfeat_config *cfg, int device, amtfeat_plan **out) {
    if (!cfg || !out) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    *out = nullptr;
    amtfeat_plan *h = new (std::nothrow) amtfeat_plan();
    if (!h) { amtfeat::set_error("out of memory"); return AMTFEAT_ERR_INVALID; }
    h->p.cfg = *cfg;
    h->p.cfg.decim_taps = nullptr;  // copied into the plan by build_plan_tables via the caller's pointer below
    h->p.device = device;
    int rc;
    try {
        amtfeat_config tmp = *cfg;
        h->p.cfg = tmp;
        rc = amtfeat::build_plan_tables(h->p);
        h->p.cfg.decim_taps = nullptr;  // never keep the caller's pointer
        h->p.cfg.n_decim_taps = 0;
        if (rc == AMTFEAT_OK && device >= 0) rc = amtfeat::upload_plan(h->p);
    } catch (const std::exception &e) {
        amtfeat::set_error(std::string("exception: ") + e.what());
        rc = AMTFEAT_ERR_INVALID;
    }
    if (rc != AMTFEAT_OK) {
        amtfeat::free_plan_device(h->p);
        delete h;
        return rc;
    }
    *out = h;
    return AMTFEAT_OK;
}

void amtfeat_plan_destroy(amtfeat_plan *plan) {
    if (!plan) return;
    amtfeat::free_plan_device(plan->p);
    delete plan;
}

int64_t amtfeat_expected_frames(const amtfeat_plan *plan, int64_t n) { return amtfeat::expected_frames(plan->p, n); }
int64_t amtfeat_output_frames(const amtfeat_plan *plan, int64_t n) { return amtfeat::output_frames(plan->p, n); }

int amtfeat_sample_range(const amtfeat_plan *plan, int64_t frames, int64_t *lo, int64_t *hi) {
    return amtfeat::sample_range(plan->p, frames, lo, hi);
}

int64_t amtfeat_num_samples_required(const amtfeat_plan *plan) {
    int64_t lo = 0, hi = 0;
    amtfeat::sample_range(plan->p, 1, &lo, &hi);
    return hi;
}

int amtfeat_times(const amtfeat_plan *plan, int64_t n, int at_start, double *out, int64_t capacity) {
    const Plan &p = plan->p;
    const amtfeat_config &c = p.cfg;
    const int64_t T = amtfeat::expected_frames(p, n);
    if (T > capacity) { amtfeat::set_error("times buffer too small"); return AMTFEAT_ERR_INVALID; }
    const double sr = c.sample_rate;
    // librosa.frames_to_time: (frames * hop).astype(int) / float(sr)   (common.py:250-256)
    for (int64_t t = 0; t < T; ++t) out[t] = (double)(t * (int64_t)c.hop_length) / sr;
    double shift = 0.0;
    bool sub = false, add = false;
    if (c.kind == AMTFEAT_VQT || c.kind == AMTFEAT_HVQT) {
        if (at_start) {  // vqt.py:217-225 (intended behaviour): floor(L_fmin / 2) / sr with the reference's own alpha
            const double alpha = std::pow(2.0, 1.0 / c.bins_per_octave) - 1;
            const double len = (1.0 / alpha) * sr / (p.harm[0].fmin + c.gamma / alpha);
            shift = std::floor(len / 2.0) / sr;
            sub = true;
        }
    } else {
        shift = (double)(c.win_length / 2) / sr;  // waveform.py:175-180
        sub = c.center && at_start;
        add = !c.center && !at_start;
    }
    if (sub) for (int64_t t = 0; t < T; ++t) out[t] -= shift;
    if (add) for (int64_t t = 0; t < T; ++t) out[t] += shift;
    return AMTFEAT_OK;
}

int amtfeat_early_ds_count(const amtfeat_plan *plan, int h) {
    if (h < 0 || h >= (int)plan->p.harm.size()) return -1;
    return plan->p.harm[h].eds_ref;
}

int amtfeat_num_channels(const amtfeat_plan *plan) { return plan->p.C; }
int amtfeat_feature_size(const amtfeat_plan *plan) { return plan->p.F; }

int amtfeat_out_shape(const amtfeat_plan *plan, int64_t n, int64_t shape[3], int *ndim) {
    const Plan &p = plan->p;
    const amtfeat_config &c = p.cfg;
    const int64_t T = amtfeat::output_frames(p, n);
    if (T < 0) { amtfeat::set_error("input too short (an uncentered frame longer than the padded signal, or a clip shorter than the early-downsampling factor of a CQT / VQT)"); return AMTFEAT_ERR_INVALID; }
    switch (c.kind) {
        case AMTFEAT_POWER: *ndim = 1; shape[0] = T; shape[1] = shape[2] = 1; break;
        case AMTFEAT_WAVEFORM: *ndim = 2; shape[0] = c.win_length; shape[1] = T; shape[2] = 1; break;
        case AMTFEAT_STFT:  // stft.py:59 returns (1, n_fft, 0) for empty audio
            *ndim = 3; shape[0] = 1; shape[1] = n == 0 ? c.n_fft : p.F; shape[2] = T; break;
        default: *ndim = 3; shape[0] = p.C; shape[1] = p.F; shape[2] = T; break;
    }
    return AMTFEAT_OK;
}

int amtfeat_plan_describe(const amtfeat_plan *plan, char *buf, size_t capacity) {
    const std::string s = amtfeat::describe(plan->p);
    if (s.size() + 1 > capacity) { amtfeat::set_error("describe buffer too small"); return AMTFEAT_ERR_INVALID; }
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return AMTFEAT_OK;
}

int amtfeat_clip_describe(const amtfeat_plan *plan, int64_t n, char *buf, size_t capacity) {
    if (!plan || !buf) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    const std::string s = amtfeat::describe_clip(plan->p, n);
    if (s.size() + 1 > capacity) { amtfeat::set_error("describe buffer too small"); return AMTFEAT_ERR_INVALID; }
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return AMTFEAT_OK;
}

size_t amtfeat_workspace_bytes(const amtfeat_plan *plan, int batch, const int64_t *n) {
    return amtfeat::workspace_bytes(plan->p, batch, n);
}

int amtfeat_launch_count(const amtfeat_plan *plan, int batch, const int64_t *n) {
    return amtfeat::launch_count(plan->p, batch, n);
}

int amtfeat_process(const amtfeat_plan *plan, const float *d_audio, const int64_t *in_offsets, const int64_t *num_samples,
                    const int64_t *out_offsets, int batch, float *d_out, void *d_workspace, size_t workspace_bytes,
                    void *stream) {
    if (!plan || !in_offsets || !num_samples || !out_offsets) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    try {
        return amtfeat::process(plan->p, d_audio, in_offsets, num_samples, out_offsets, batch, d_out, d_workspace,
                                workspace_bytes, stream);
    } catch (const std::exception &e) {
        amtfeat::set_error(std::string("exception: ") + e.what());
        return AMTFEAT_ERR_INVALID;
    }
}

int amtfeat_process_raw(const amtfeat_plan *plan, const float *d_audio, const int64_t *in_offsets, const int64_t *num_samples,
                        const int64_t *out_offsets, int batch, float *d_out, void *d_workspace, size_t workspace_bytes,
                        void *stream) {
    if (!plan || !in_offsets || !num_samples || !out_offsets) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    try {
        return amtfeat::process(plan->p, d_audio, in_offsets, num_samples, out_offsets, batch, d_out, d_workspace,
                                workspace_bytes, stream, /*defer_epilogue=*/true);
    } catch (const std::exception &e) {
        amtfeat::set_error(std::string("exception: ") + e.what());
        return AMTFEAT_ERR_INVALID;
    }
}

int amtfeat_range_reference(const amtfeat_plan *plan, const float *d_block, int64_t frames, int64_t t_begin, int64_t t_end,
                            float *d_ref, void *stream) {
    if (!plan || !d_block || !d_ref) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    return amtfeat::range_reference(plan->p, d_block, frames, t_begin, t_end, d_ref, stream);
}

int amtfeat_range_finish(const amtfeat_plan *plan, const float *d_block, int64_t frames, int64_t t_begin, int64_t t_end,
                         const float *d_ref, float *d_dst, int64_t dst_frames, int64_t t_dst, void *stream) {
    if (!plan || !d_block || !d_dst || !d_ref) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    return amtfeat::range_finish(plan->p, d_block, frames, t_begin, t_end, d_ref, d_dst, dst_frames, t_dst, stream);
}

int amtfeat_process_host(const amtfeat_plan *plan, const float *h_audio, const int64_t *in_offsets,
                         const int64_t *num_samples, const int64_t *out_offsets, int batch, float *h_out,
                         int64_t audio_elems, int64_t out_elems, float *d_audio, float *d_out, void *d_workspace,
                         size_t workspace_bytes, void *stream) {
    if (!plan) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    if (plan->p.device < 0) { amtfeat::set_error("host-only plan: no CUDA device (there is no CPU compute path)"); return AMTFEAT_ERR_NO_DEVICE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemcpyAsync(d_audio, h_audio, (size_t)audio_elems * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { amtfeat::set_error(std::string("H2D: ") + cudaGetErrorString(e)); return AMTFEAT_ERR_CUDA; }
    int rc = amtfeat_process(plan, d_audio, in_offsets, num_samples, out_offsets, batch, d_out, d_workspace, workspace_bytes, stream);
    if (rc != AMTFEAT_OK) return rc;
    e = cudaMemcpyAsync(h_out, d_out, (size_t)out_elems * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { amtfeat::set_error(std::string("D2H: ") + cudaGetErrorString(e)); return AMTFEAT_ERR_CUDA; }
    return AMTFEAT_OK;
}

// ---- pipelined host executor -----------------------------------------------------------------------------------
}  // extern "C"

struct amtfeat_pipeline {
    int device = 0, nslots = 0;
    int64_t max_audio = 0, max_out = 0;
    size_t max_ws = 0;
    cudaStream_t s_h2d = nullptr, s_compute = nullptr, s_d2h = nullptr;
    struct Slot {
        float *d_audio = nullptr, *d_out = nullptr;
        void *d_ws = nullptr;
        cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr;
        int64_t ticket = -1;
    };
    std::vector<Slot> slots;
    int64_t next_ticket = 0;
    bool fused_epilogue = true;   // dB epilogue fused into the download (AMTFEAT_PIPE_FUSED=0: in-place pass + copy)
};

#define PIPE_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            amtfeat::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));       \
            return AMTFEAT_ERR_CUDA;                                                      \
        }                                                                                 \
    } while (0)

extern "C" {

void amtfeat_pipeline_destroy(amtfeat_pipeline *pipe) {
    if (!pipe) return;
    amtfeat::DeviceGuard guard(pipe->device);
    if (pipe->s_d2h) cudaStreamSynchronize(pipe->s_d2h);
    for (auto &s : pipe->slots) {
        if (s.d_audio) cudaFree(s.d_audio);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_ws) cudaFree(s.d_ws);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
        if (s.computed) cudaEventDestroy(s.computed);
        if (s.downloaded) cudaEventDestroy(s.downloaded);
    }
    if (pipe->s_h2d) cudaStreamDestroy(pipe->s_h2d);
    if (pipe->s_compute) cudaStreamDestroy(pipe->s_compute);
    if (pipe->s_d2h) cudaStreamDestroy(pipe->s_d2h);
    delete pipe;
}

static int pipeline_init(amtfeat_pipeline *p) {
    amtfeat::DeviceGuard guard(p->device);
    PIPE_CUDA(guard.status);
    PIPE_CUDA(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
    PIPE_CUDA(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
    {
        // the fused epilogue + download kernel must not queue behind the compute grids of the next batch: highest priority
        int prio_lo = 0, prio_hi = 0;
        PIPE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const char *env = std::getenv("AMTFEAT_PIPE_PRIO");
        PIPE_CUDA(cudaStreamCreateWithPriority(&p->s_d2h, cudaStreamNonBlocking, (env && std::string(env) == "0") ? prio_lo : prio_hi));
    }
    for (auto &s : p->slots) {
        PIPE_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_audio), (size_t)std::max<int64_t>(p->max_audio, 4) * sizeof(float)));
        PIPE_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_out), (size_t)std::max<int64_t>(p->max_out, 4) * sizeof(float)));
        PIPE_CUDA(cudaMalloc(&s.d_ws, std::max<size_t>(p->max_ws, 256)));
        PIPE_CUDA(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
        PIPE_CUDA(cudaEventCreateWithFlags(&s.computed, cudaEventDisableTiming));
        PIPE_CUDA(cudaEventCreateWithFlags(&s.downloaded, cudaEventDisableTiming));
    }
    return AMTFEAT_OK;
}

int amtfeat_pipeline_create(int device, int nslots, int64_t max_audio_elems, int64_t max_out_elems, size_t max_workspace_bytes,
                            amtfeat_pipeline **out) {
    if (!out || nslots < 1 || nslots > 64 || max_audio_elems < 0 || max_out_elems < 0) { amtfeat::set_error("invalid pipeline arguments"); return AMTFEAT_ERR_INVALID; }
    *out = nullptr;
    if (device < 0) { amtfeat::set_error("a pipeline needs a CUDA device (there is no CPU compute path)"); return AMTFEAT_ERR_NO_DEVICE; }
    amtfeat_pipeline *p = new (std::nothrow) amtfeat_pipeline();
    if (!p) { amtfeat::set_error("out of memory"); return AMTFEAT_ERR_INVALID; }
    p->device = device; p->nslots = nslots; p->max_audio = max_audio_elems; p->max_out = max_out_elems; p->max_ws = max_workspace_bytes;
    p->slots.resize(nslots);
    {
        const char *env = std::getenv("AMTFEAT_PIPE_FUSED");
        p->fused_epilogue = !(env && std::string(env) == "0");
    }
    const int rc = pipeline_init(p);
    if (rc != AMTFEAT_OK) { amtfeat_pipeline_destroy(p); return rc; }
    *out = p;
    return AMTFEAT_OK;
}

int amtfeat_pipeline_submit(amtfeat_pipeline *pipe, const amtfeat_plan *plan, const float *h_audio, const int64_t *in_offsets,
                            const int64_t *num_samples, const int64_t *out_offsets, int batch, float *h_out, int64_t audio_elems,
                            int64_t out_elems, int64_t *ticket) {
    if (!pipe || !plan) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    if (plan->p.device != pipe->device) { amtfeat::set_error("plan and pipeline live on different devices"); return AMTFEAT_ERR_INVALID; }
    if (audio_elems > pipe->max_audio || out_elems > pipe->max_out) { amtfeat::set_error("batch larger than the pipeline's staging buffers"); return AMTFEAT_ERR_INVALID; }
    const size_t ws = amtfeat::workspace_bytes(plan->p, batch, num_samples);
    if (ws > pipe->max_ws) { amtfeat::set_error("batch needs a larger workspace than the pipeline was created with"); return AMTFEAT_ERR_WORKSPACE; }
    amtfeat::DeviceGuard guard(pipe->device);
    PIPE_CUDA(guard.status);
    amtfeat_pipeline::Slot &s = pipe->slots[pipe->next_ticket % pipe->nslots];
    // upload: the kernels that last read this slot's audio must be done
    PIPE_CUDA(cudaStreamWaitEvent(pipe->s_h2d, s.computed, 0));
    PIPE_CUDA(cudaMemcpyAsync(s.d_audio, h_audio, (size_t)audio_elems * sizeof(float), cudaMemcpyHostToDevice, pipe->s_h2d));
    PIPE_CUDA(cudaEventRecord(s.uploaded, pipe->s_h2d));
    // compute: after the upload, and after the previous download of this slot's output
    PIPE_CUDA(cudaStreamWaitEvent(pipe->s_compute, s.uploaded, 0));
    PIPE_CUDA(cudaStreamWaitEvent(pipe->s_compute, s.downloaded, 0));
    // dB features: the (x - max) / 80 + 1 pass is fused into the download -- one kernel on the download stream reads the raw log
    // values from HBM once and stores the finished features straight into the caller's pinned (mapped) host buffer, instead of an
    // in-place pass over HBM followed by a copy.  AMTFEAT_PIPE_FUSED=0 keeps the two-step form (A/B).
    const bool fused = pipe->fused_epilogue && amtfeat::has_db_epilogue(plan->p);
    int rc, maxT = 0;
    try {
        rc = amtfeat::process(plan->p, s.d_audio, in_offsets, num_samples, out_offsets, batch, s.d_out, s.d_ws, pipe->max_ws, pipe->s_compute,
                              fused, &maxT);
    } catch (const std::exception &e) {
        amtfeat::set_error(std::string("exception: ") + e.what());
        rc = AMTFEAT_ERR_INVALID;
    }
    if (rc != AMTFEAT_OK) return rc;
    PIPE_CUDA(cudaEventRecord(s.computed, pipe->s_compute));
    // download
    PIPE_CUDA(cudaStreamWaitEvent(pipe->s_d2h, s.computed, 0));
    if (fused) {
        rc = amtfeat::epilogue_out(plan->p, s.d_out, s.d_ws, batch, maxT, h_out, true, pipe->s_d2h);
        if (rc != AMTFEAT_OK) return rc;
    } else {
        PIPE_CUDA(cudaMemcpyAsync(h_out, s.d_out, (size_t)out_elems * sizeof(float), cudaMemcpyDeviceToHost, pipe->s_d2h));
    }
    PIPE_CUDA(cudaEventRecord(s.downloaded, pipe->s_d2h));
    s.ticket = pipe->next_ticket;
    if (ticket) *ticket = pipe->next_ticket;
    ++pipe->next_ticket;
    return AMTFEAT_OK;
}

int amtfeat_pipeline_wait(amtfeat_pipeline *pipe, int64_t ticket) {
    if (!pipe) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    amtfeat::DeviceGuard guard(pipe->device);
    PIPE_CUDA(guard.status);
    if (ticket < 0 || ticket + pipe->nslots < pipe->next_ticket) {   // everything (or a ticket whose slot was already recycled)
        PIPE_CUDA(cudaStreamSynchronize(pipe->s_d2h));
        return AMTFEAT_OK;
    }
    if (ticket >= pipe->next_ticket) { amtfeat::set_error("unknown ticket"); return AMTFEAT_ERR_INVALID; }
    PIPE_CUDA(cudaEventSynchronize(pipe->slots[ticket % pipe->nslots].downloaded));
    return AMTFEAT_OK;
}

int64_t amtfeat_framify_hops(int64_t num_frames, int win_length, int hop_length, int pad) {
    if (win_length <= 0 || hop_length <= 0 || num_frames < 0) return -1;
    const int64_t pad_length = win_length / 2;
    const int64_t padded = pad ? num_frames + 2 * pad_length : (num_frames > win_length ? num_frames : win_length);
    return (padded - 2 * pad_length) / hop_length;   // tools/utils.py:2971
}

int amtfeat_framify(const float *d_in, int64_t rows, int64_t num_frames, int win_length, int hop_length, int pad, float *d_out,
                    void *stream) {
    const int64_t hops = amtfeat_framify_hops(num_frames, win_length, hop_length, pad);
    if (hops < 0 || rows < 0) { amtfeat::set_error("invalid framify arguments"); return AMTFEAT_ERR_INVALID; }
    const int64_t pad_length = win_length / 2;
    const int64_t padded = pad ? num_frames + 2 * pad_length : (num_frames > win_length ? num_frames : win_length);
    const int64_t lpad = (padded - num_frames) / 2;   // librosa.util.pad_center
    return amtfeat::framify(d_in, rows, num_frames, win_length, hop_length, lpad, hops, d_out, stream);
}

int amtfeat_profile_enable(amtfeat_plan *plan, int enable) {
    if (!plan) return AMTFEAT_ERR_INVALID;
    plan->p.prof_enabled = enable != 0;
    return AMTFEAT_OK;
}

int amtfeat_profile_read(amtfeat_plan *plan, char *buf, size_t capacity) {
    if (!plan) return AMTFEAT_ERR_INVALID;
    std::string s;
    int rc = amtfeat::profile_read(plan->p, s);
    if (rc != AMTFEAT_OK) return rc;
    if (s.size() + 1 > capacity) { amtfeat::set_error("profile buffer too small"); return AMTFEAT_ERR_INVALID; }
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return AMTFEAT_OK;
}

}  // extern "C"
