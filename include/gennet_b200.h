/*
 * gennet_b200 -- C ABI of the B200 (sm_100a) hot path of hagabbar/GenNet.
 *
 * The reference has no FFI of its own (it is pure Python on Keras/TensorFlow/NumPy),
 * so each entry point names the reference call site (file:line, relative to the
 * reference checkout) whose arithmetic it replaces.  A maintainer binds these
 * with ctypes; gennet_b200/_lib.py is that binding.
 *
 * Conventions
 *   - every pointer is caller-owned DEVICE memory unless the name ends in _host;
 *   - tensors are dense row-major; activations are channels-last (NLC / NHWC) as in
 *     Keras; weights use the Keras layouts (Dense (in,out); Conv1D (k,Cin,Cout));
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and
 *     nothing synchronises the device;
 *   - return value: GN_OK, or a negative GN_ERR_* with a message in gn_last_error()
 *     (thread-local).  The library never falls back to the CPU.
 */
#ifndef GENNET_B200_H
#define GENNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GN_OK 0
#define GN_ERR_ARG (-1)         /* invalid argument / shape */
#define GN_ERR_UNSUPPORTED (-2) /* valid but not implemented for this shape */
#define GN_ERR_CUDA (-3)        /* CUDA runtime / driver error */

/* activation codes (Keras Activation / LeakyReLU / ReLU(max_value), bbhMahoGANy.py:238,363,400,440,495) */
#define GN_ACT_NONE 0
#define GN_ACT_RELU 1
#define GN_ACT_TANH 2
#define GN_ACT_SIGMOID 3
#define GN_ACT_LEAKY 4    /* param = alpha */
#define GN_ACT_RELU_MAX 5 /* param = max_value */
#define GN_ACT_ELU 6      /* alpha = 1 (2_model_version/weight_version/subtract_model.py:215) */

/* dropout-family codes (Keras Dropout / GaussianDropout / GaussianNoise) */
#define GN_NOISE_DROPOUT 0   /* r = keep mask (0/1):  y = x*r/(1-rate)              */
#define GN_NOISE_GDROPOUT 1  /* r = N(0,1):           y = x*(1+r*sqrt(rate/(1-rate))) */
#define GN_NOISE_GNOISE 2    /* r = N(0,1):           y = x + r*rate (rate = stddev) */

/* loss codes (Keras 'binary_crossentropy', 'mean_squared_error', chisquare_Loss bbhMahoGANy.py:146-162) */
#define GN_LOSS_BCE 0
#define GN_LOSS_MSE 1
#define GN_LOSS_CHISQ 2 /* param = n_sig */

const char* gn_last_error(void);
int gn_version(void);
/* 1 if the visible device is sm_100 (B200) and the kernels can run, else 0 with gn_last_error() set */
int gn_device_ok(void);

/* ---------------------------------------------------------------------------
 * Sample synthesis  (BBH_version/gw_template_maker.py)
 * ------------------------------------------------------------------------- */

typedef struct gn_fft_plan gn_fft_plan;

/* Twiddle tables for real FFTs of length N (power of two, 512 <= N <= 32768). */
int gn_fft_plan_create(int N, gn_fft_plan** plan);
int gn_fft_plan_destroy(gn_fft_plan* plan);

/* whiten_data(x, T, fs, psd, 'td')  gw_template_maker.py:243-286, batched, + crop (:695) + scale (:813-814):
 *   y[b, j] = scale * irfft( rfft(window * x[b]) * weights )[crop_lo + j],  j < crop_len
 * x (batch,N) f32; window (N) f32 = tukey(N,1/8); weights (N/2+1) f32 = sqrt(2/(S*fs)), 0 where S<=0, [0]=0.
 * Each call first rebuilds, on `stream`, a per-bin coefficient table from `weights`, `window` and `scale` in scratch
 * owned by the plan (slots keyed by the weights and window pointers and the scale, guarded by events): a plan may be used from several streams
 * and host threads; as with any asynchronous call, weights/window must not be overwritten while a call that reads
 * them is still in flight.  The same holds for gn_synth_f32. */
int gn_whiten_td_f32(const gn_fft_plan* plan, const float* x, const float* window, const float* weights,
                     float* y, int batch, int crop_lo, int crop_len, float scale, void* stream);

/* scale * irfft(xf * weights) rolled by `roll` samples (np.roll semantics), batched.
 * Covers whiten_data(...,'fd') + np.fft.irfft + np.roll of gen_bbh (:518-522, roll=-fs),
 * main()'s event whitening (:774-777) and gen_noise (:184-191; weights=amp, scale=N*df).
 * xf (batch, N/2+1) interleaved complex f32; weights may be NULL (all ones). Imag of the DC and
 * Nyquist bins is ignored as numpy's irfft does; drop_dc!=0 zeroes the DC bin (:189-190,:279). */
int gn_irfft_f32(const gn_fft_plan* plan, const float* xf, const float* weights, float* y, int batch,
                 float scale, int roll, int drop_dc, void* stream);

/* Fused per-batch training-sample synthesis (sim_data noise branch :685-691 as restated in SURVEY a7):
 *   n      = fs * irfft(amp * (normals_re + i normals_im)), DC = 0            (gen_noise :184-191)
 *   out[b] = out_scale * whiten_td(n + templates[tidx[b]])[crop_lo : crop_lo+crop_len]
 * normals (batch,2,N/2+1) f32 standard normals (re row then im row, the order of :187-188), or NULL to draw
 * them in-kernel with Philox4x32-10 keyed by (seed, sample_offset+b); amp (N/2+1) = sqrt(T*S/4);
 * templates (n_templates,N) f32 or NULL (noise only); tidx (batch) int32 or NULL (template b). */
int gn_synth_f32(const gn_fft_plan* plan, const float* normals, const float* amp, const float* templates,
                 const int* tidx, const float* window, const float* weights, float* out, int batch,
                 int n_templates, int crop_lo, int crop_len, float noise_scale, float out_scale,
                 uint64_t seed, uint64_t sample_offset, void* stream);

/* Tail of gen_bbh (gw_template_maker.py:528-571) + crop (:695) + norm (:813-814), batched:
 *   ref = argmax(hp^2+hc^2); ht = Fp*hp + Fc*hc; ts[:len] = ht[ref-idx-lead:] (Python slice semantics,
 *   negative start wraps); ts *= win; out = scale * ts[crop_lo : crop_lo+crop_len].
 * hp, hc (batch,N) whitened rolled polarisations (gn_irfft_f32 with roll=-fs); Fp, Fc (batch) antenna factors
 * (pylal.antenna.response is not restated: the caller supplies them); ref_idx (batch) int32 out or NULL. */
int gn_bbh_assemble_f32(const float* hp, const float* hc, const float* Fp, const float* Fc, const int* idx,
                        int lead, const float* win, float* out, int* ref_idx, int batch, int N, int crop_lo,
                        int crop_len, float scale, void* stream);

/* mean and population std of n floats (np.std, gw_template_maker.py:782); out_host-free: out (2) f32 on device */
int gn_mean_std_f32(const float* x, long long n, float* out, void* stream);

/* x[r, :] += sigma * normals[r, :] for r < rows  (on-the-fly noise, bbhMahoGANy.py:1161,1277,1030) */
int gn_add_scaled_f32(float* x, const float* r, float sigma, long long n, void* stream);

/* sine-Gaussian bursts, tests/burstMahoGANy.py:76-98: out[i,j] = amp*sin(2*pi*freq*(t_j-t0_i)+phi)*exp(-(t_j-t0_i)^2/tau_i^2)
 * pars (n,2) f64 = (t0, tau): kept in double because the phase 2*pi*f*(t-t0) reaches ~300 rad */
int gn_burst_waveforms_f32(const double* pars, float* out, int n, int N, float amp, float freq, float dt,
                           float phi, void* stream);

/* ---------------------------------------------------------------------------
 * Network layers (Keras semantics; bbhMahoGANy.py:212-539, tests/burstMahoGANy.py:127-423,
 * train_on_wvf_version/nn.py:72-106)
 * ------------------------------------------------------------------------- */

/* Conv1D, NLC, kernel (k,Cin,Cout), zero padding pad_left (TF 'SAME' rule computed by the caller), stride s.
 * up = 1, or 2 to read the input through a fused UpSampling1D(2) (bbhMahoGANy.py:249-250,258-259):
 * x is then the (B, L/2, Cin) tensor before upsampling and L the upsampled length.
 * y = act(conv(x) + bias). */
int gn_conv1d_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int L, int Cin,
                      int Lout, int Cout, int k, int stride, int pad_left, int up, int act, float act_param,
                      void* stream);
/* dx (B, L/up, Cin) = conv-transpose(dy, w); overwrites dx */
int gn_conv1d_dgrad_f32(const float* dy, const float* w, float* dx, int B, int L, int Cin, int Lout, int Cout,
                        int k, int stride, int pad_left, int up, void* stream);
/* dw (k,Cin,Cout) and db (Cout) OVERWRITTEN with the batch gradient (db may be NULL) */
int gn_conv1d_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int L, int Cin, int Lout,
                        int Cout, int k, int stride, int pad_left, int up, void* stream);

/* ---- tensor-core (tcgen05 + TMA) Conv1D, bf16 activations/weights, fp32 accumulation --------------------
 * Same semantics as gn_conv1d_*_f32 for k <= 16, stride 1 or 2, Cin % 64 == 0 and Cout % 64 == 0 (the first,
 * Cin=1, and last, Cout=1, layers of the reference networks are bandwidth-bound and stay on the SIMT kernels).
 * All activation / weight pointers are bf16 device memory, 16-byte aligned.
 *   gn_conv_w_to_bf16 : w f32 (k,Cin,Cout) -> wk bf16 (k,Cin,Cout) [dgrad operand] and wt bf16 (k,Cout,Cin) [fwd operand]
 *   fwd   : y = act(conv(x, w) + bias)                      x (B,L,Cin), y (B,Lout,Cout)
 *   dgrad : dx = act'(x_in) * conv_transpose(dy, w)         x_in = this conv's input (or NULL): fuses the backward
 *           of the activation layer that produced x_in (in_act = its GN_ACT_* code).  dx_colsum (f32 (Cin),
 *           OVERWRITTEN, or NULL): per-channel sum of dx over (b, l) = the bias gradient of the convolution
 *           that produced x_in (dL/db = sum of its pre-activation output gradient), taken from the epilogue's
 *           shared-memory slabs so that layer's wgrad call can pass db = NULL and skip a pass over dx
 *   wgrad : dw f32 (k,Cin,Cout), db f32 (Cout) OVERWRITTEN  (needs Cin % 128 == 0, or Cin == 64 and Cout % 128 == 0) */
int gn_conv_w_to_bf16(const float* w, void* wk, void* wt, int k, int Cin, int Cout, void* stream);
int gn_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream);
int gn_cast_bf16_to_f32(const void* x, float* y, long long n, void* stream);
int gn_conv1d_fwd_bf16(const void* x, const void* wt, const float* bias, void* y, int B, int L, int Cin, int Lout,
                       int Cout, int k, int stride, int pad_left, int act, float act_param, void* stream);
int gn_conv1d_dgrad_bf16(const void* dy, const void* wk, const void* x_in, void* dx, float* dx_colsum, int B, int L,
                         int Cin, int Lout, int Cout, int k, int stride, int pad_left, int in_act, float in_act_param,
                         void* stream);
int gn_conv1d_wgrad_bf16(const void* x, const void* dy, float* dw, float* db, int B, int L, int Cin, int Lout,
                         int Cout, int k, int stride, int pad_left, void* stream);

/* The same bandwidth-bound kernels with FLOAT32 activations (float32 and split-operand modes, whose activations are
 * float32): identical arguments and semantics, every `void*` activation / gradient tensor above is `float*` here.
 * The float32 smallcin entry points also take ANY filter count up to 1024 and up to 16 taps (k * Cin * Cout weights must
 * fit 48 KB of shared memory; dgrad: k * Cin <= 32): Conv1D(50, 16) on (out_dim, 1) of the 2_model_version
 * discriminators (no_mode_collapse_network.py:117, subtract_model.py:134), Conv1D(25, 5) on (8192, 1) of
 * train_on_wvf_version/nn.py:95. */
int gn_conv1d_smallcin_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int L, int Cin,
                               int Lout, int Cout, int k, int stride, int pad_left, int act, float act_param,
                               void* stream);
int gn_conv1d_smallcin_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int L, int Cin, int Lout,
                                 int Cout, int k, int stride, int pad_left, void* stream);
int gn_conv1d_smallcin_dgrad_f32(const float* dy, const float* w, float* dx, int B, int L, int Cin, int Lout, int Cout,
                                 int k, int stride, int pad_left, void* stream);
int gn_conv1d_cout1_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int L, int Cin, int Lout,
                            int k, int pad_left, void* stream);
int gn_conv1d_cout1_dgrad_f32(const float* dy, const float* w, float* dx, int B, int L, int Cin, int Lout, int k,
                              int pad_left, void* stream);
int gn_conv1d_cout1_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int L, int Cin, int Lout,
                              int k, int pad_left, void* stream);
int gn_dense_small_fwd_f32(const float* x, const float* w, const float* bias, float* y, int M, int K, int N, int act,
                           float act_param, void* stream);
int gn_dense_small_dgrad_f32(const float* dy, const float* w, const float* x_in, float* dx, float* dx_colsum,
                             int colsum_channels, int M, int K, int N, int in_act, float in_act_param, void* stream);
int gn_dense_small_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int M, int K, int N, void* stream);

/* ---- tensor-core Conv1D at FLOAT32 accuracy: split-bf16 operands ("bf16x3") -------------------------------
 * The reference's Conv1D layers compute in float32 (bbhMahoGANy.py:250-292,362-395 through cuDNN/Eigen).  Here a
 * float32 tensor t is carried as nc bf16 planes, t = t0 + t1 (+ t2) with t0 = bf16(t), t1 = bf16(t - t0), ...,
 * stored plane-major as (nc, B, L, C), and a product is accumulated in fp32 tensor memory as the sum of the plane
 * products t_i * w_j with i + j < nc (nc = 3: six tcgen05.mma per K step, float32-class accuracy; nc = 2: three,
 * ~2^-16; nc = 1 is plain bf16).  Same geometry limits as gn_conv1d_*_bf16.
 *   gn_split_f32_bf16    : x f32 (n) -> planes bf16 (nc, n);  n % 8 == 0
 *   gn_conv_w_split_bf16 : w f32 (k,Cin,Cout) -> wk planes (nc,k,Cin,Cout) [dgrad operand], wt planes (nc,k,Cout,Cin) [fwd]
 *   fwd   : y f32 (B,Lout,Cout) and / or ys planes (nc,B,Lout,Cout) = act(conv(x, w) + bias); either may be NULL
 *   dgrad : dx f32 (B,L,Cin) and / or dxs planes = act'(x_in) * conv_transpose(dy, w); x_in f32 = this conv's input
 *           or NULL; dx_colsum f32 (Cin) OVERWRITTEN or NULL (bias gradient of the convolution that produced x_in)
 *   wgrad : dw f32 (k,Cin,Cout) OVERWRITTEN from the x and dy planes; db f32 (Cout) OVERWRITTEN from the float32 dy
 *           (NULL: skipped) */
int gn_split_f32_bf16(const float* x, void* planes, long long n, int nc, void* stream);
int gn_conv_w_split_bf16(const float* w, void* wk, void* wt, int k, int Cin, int Cout, int nc, void* stream);
int gn_conv1d_fwd_bf16x3(const void* xs, const void* wts, const float* bias, float* y, void* ys, int B, int L, int Cin,
                         int Lout, int Cout, int k, int stride, int pad_left, int act, float act_param, int nc,
                         void* stream);
int gn_conv1d_dgrad_bf16x3(const void* dys, const void* wks, const float* x_in, float* dx, void* dxs, float* dx_colsum,
                           int B, int L, int Cin, int Lout, int Cout, int k, int stride, int pad_left, int in_act,
                           float in_act_param, int nc, void* stream);
int gn_conv1d_wgrad_bf16x3(const void* xs, const void* dys, const float* dy, float* dw, float* db, int B, int L, int Cin,
                           int Lout, int Cout, int k, int stride, int pad_left, int nc, void* stream);

/* Dense layers on the same split-operand kernels (Dense(100 -> 128 n_pix) of the generator, bbhMahoGANy.py:234;
 * Dense(16128 -> 1024) of the burst discriminator, burstMahoGANy.py:351; the out x out layers of 2_model_version):
 * y (M,N) = act(x (M,K) w (K,N) + bias) is a one-tap convolution over a single sample whose positions are the M batch
 * rows.  Kp = K rounded up to the channel tile (64, or 128 for the weight gradient) exists in the PLANES only, zero filled.
 *   gn_split_pad_f32_bf16 : x f32 (rows,K) -> planes bf16 (nc,rows,Kp)
 *   gn_dense_w_split_bf16 : w f32 (K,N) -> wk planes (nc,Kp,N) [dgrad operand], wt planes (nc,N,Kp) [fwd operand]
 *   fwd   : y f32 (M,N) and / or ys planes (nc,M,N); N % 64 == 0, Kp % 64 == 0
 *   dgrad : dx f32 (M,K) = act'(x_in) * dy w^T; needs K % 64 == 0 (no padding); dx_colsum as for the convolution
 *   wgrad : dw f32 (K,N) OVERWRITTEN (only the K real rows are written), db f32 (N) from the float32 dy;
 *           Kp % 128 == 0, or Kp == 64 with N % 128 == 0 */
int gn_split_pad_f32_bf16(const float* x, void* planes, long long rows, int K, int Kp, int nc, void* stream);
int gn_dense_w_split_bf16(const float* w, void* wk, void* wt, int K, int Kp, int N, int nc, void* stream);
int gn_dense_fwd_bf16x3(const void* xs, const void* wts, const float* bias, float* y, void* ys, int M, int Kp, int N,
                        int act, float act_param, int nc, void* stream);
int gn_dense_dgrad_bf16x3(const void* dys, const void* wks, const float* x_in, float* dx, float* dx_colsum, int M, int K,
                          int N, int in_act, float in_act_param, int nc, void* stream);
int gn_dense_wgrad_bf16x3(const void* xs, const void* dys, const float* dy, float* dw, float* db, int M, int K, int N,
                          int Kp, int nc, void* stream);

/* ---- the same split-operand kernels with SCALED FP16 PAIRS ("f16x2"): half the tensor-core work of bf16x3 -------
 * A float32 tensor t with max |t| = amax is carried as two fp16 planes (2, ...) plus the device scalar amax:
 *     t = (T0 + 2^-11 T1) / s,  s = 2^(14 - floor(log2 amax)),  T0 = fp16(t s),  T1 = fp16((t s - T0) 2^11)
 * i.e. 22-23 significant bits relative to the TENSOR's largest element (elements below 2^-28 amax lose bits), and a
 * product needs three plane products (T0 W0 | T0 W1 + T1 W0; the dropped term is <= 2^-22 |t||w|); the kernels undo
 * both scales in the epilogue.  Results are float32; geometry, outputs and OVERWRITE rules as the _bf16x3 entry points.
 *   gn_amax_f32            : amax[0] = max |x| (device scalar)
 *   gn_split_f32_f16x2     : x f32 (n) -> planes fp16 (2, n), amax[0] = max |x| (have_amax != 0: amax[0] already holds
 *                            it, e.g. from the y_amax / dx_amax output of the producing convolution)
 *   gn_conv_w_split_f16x2, gn_split_pad_f32_f16x2, gn_dense_w_split_f16x2 : as their _bf16 counterparts, plus amax
 *   fwd / dgrad            : y_amax / dx_amax (device scalar, OVERWRITTEN, may be NULL) = max |result| */
int gn_amax_f32(const float* x, long long n, float* amax, void* stream);
int gn_split_f32_f16x2(const float* x, void* planes, float* amax, int have_amax, long long n, void* stream);
int gn_conv_w_split_f16x2(const float* w, void* wk, void* wt, float* amax, int k, int Cin, int Cout, void* stream);
/* the split of a gradient tensor dy f32 (rows, C) that also leaves colsum f32 (C) OVERWRITTEN = its per-channel column
 * sums, i.e. the bias gradient of the layer it belongs to (pass db = NULL to the weight-gradient call then);
 * C % 8 == 0 and C / 8 divides 256 or is a multiple of it */
int gn_split_colsum_f32_f16x2(const float* x, void* planes, float* amax, int have_amax, long long rows, int C, float* colsum,
                              void* stream);
int gn_conv1d_fwd_f16x2(const void* xs, const float* x_amax, const void* wts, const float* w_amax, const float* bias,
                        float* y, float* y_amax, int B, int L, int Cin, int Lout, int Cout, int k, int stride,
                        int pad_left, int act, float act_param, void* stream);
/* forward with the BatchNormalization statistics of the layer that follows: y_sums (2 * Cout doubles, OVERWRITTEN) =
 * per-channel (sum y, sum y^2) of the stored result, taken from the accumulators in the epilogue (Cout <= 1024) */
int gn_conv1d_fwd_stats_f16x2(const void* xs, const float* x_amax, const void* wts, const float* w_amax, const float* bias,
                              float* y, double* y_sums, int B, int L, int Cin, int Lout, int Cout, int k, int stride,
                              int pad_left, int act, float act_param, void* stream);
int gn_conv1d_dgrad_f16x2(const void* dys, const float* dy_amax, const void* wks, const float* w_amax, const float* x_in,
                          float* dx, float* dx_colsum, float* dx_amax, int B, int L, int Cin, int Lout, int Cout, int k,
                          int stride, int pad_left, int in_act, float in_act_param, void* stream);
int gn_conv1d_wgrad_f16x2(const void* xs, const float* x_amax, const void* dys, const float* dy_amax, const float* dy,
                          float* dw, float* db, int B, int L, int Cin, int Lout, int Cout, int k, int stride, int pad_left,
                          void* stream);
int gn_split_pad_f32_f16x2(const float* x, void* planes, float* amax, long long rows, int K, int Kp, void* stream);
int gn_dense_w_split_f16x2(const float* w, void* wk, void* wt, float* amax, int K, int Kp, int N, void* stream);
int gn_dense_fwd_f16x2(const void* xs, const float* x_amax, const void* wts, const float* w_amax, const float* bias, float* y,
                       int M, int Kp, int N, int act, float act_param, void* stream);
int gn_dense_dgrad_f16x2(const void* dys, const float* dy_amax, const void* wks, const float* w_amax, const float* x_in,
                         float* dx, float* dx_colsum, int M, int K, int N, int in_act, float in_act_param, void* stream);
int gn_dense_wgrad_f16x2(const void* xs, const float* x_amax, const void* dys, const float* dy_amax, const float* dy,
                         float* dw, float* db, int M, int K, int N, int Kp, void* stream);

/* Bandwidth-bound companions of the bf16 path.
 *   smallcin fwd  : first convolution of a network, Cin in {1,2}: x f32 (B,L,Cin) -> y bf16 (B,Lout,Cout), bias+act fused
 *   smallcin wgrad: dw f32 (k,Cin,Cout), db f32 (Cout) OVERWRITTEN from x f32 and dy bf16 (k <= 5, Cout in {8,16,32,64}
 *                   or a multiple of 128)
 *   dense_small_* : Dense with N <= 4 outputs over bf16 features (K % 8 == 0): fwd y f32 (M,N); dgrad dx bf16 (M,K)
 *                   = act'(x_in) * dy w^T (x_in = the layer's input or NULL), dx_colsum f32 (colsum_channels)
 *                   OVERWRITTEN or NULL: sum of dx over rows and over features k with equal k % colsum_channels
 *                   (bias gradient of the convolution whose flattened (L, C) output feeds this layer);
 *                   wgrad dw f32 (K,N), db f32 (N) OVERWRITTEN
 *   gn_act_bwd_bf16: dx = dy * act'(y) on bf16 tensors */
int gn_conv1d_smallcin_fwd_bf16(const float* x, const float* w, const float* bias, void* y, int B, int L, int Cin,
                                int Lout, int Cout, int k, int stride, int pad_left, int act, float act_param,
                                void* stream);
int gn_conv1d_smallcin_wgrad_bf16(const float* x, const void* dy, float* dw, float* db, int B, int L, int Cin, int Lout,
                                  int Cout, int k, int stride, int pad_left, void* stream);
/* data gradient of a Cin in {1,2} convolution: dx f32 (B,L,Cin) OVERWRITTEN from dy bf16 (B,Lout,Cout) and w f32
 * (k,Cin,Cout), k <= 5 (generator step through the frozen discriminator's first layer, bbhMahoGANy.py:1296) */
int gn_conv1d_smallcin_dgrad_bf16(const void* dy, const float* w, float* dx, int B, int L, int Cin, int Lout, int Cout,
                                  int k, int stride, int pad_left, void* stream);
/* last convolution of the generator (Cout = 1, stride 1, k <= 5; bbhMahoGANy.py:291): x bf16 (B,L,Cin) -> y f32 (B,Lout)
 * OVERWRITTEN (bias added, no activation); dgrad dy f32 (B,Lout) -> dx bf16 (B,L,Cin); wgrad dw f32 (k,Cin), db f32 (1)
 * OVERWRITTEN.  One pass over the big operand each. */
int gn_conv1d_cout1_fwd_bf16(const void* x, const float* w, const float* bias, float* y, int B, int L, int Cin, int Lout,
                             int k, int pad_left, void* stream);
int gn_conv1d_cout1_dgrad_bf16(const float* dy, const float* w, void* dx, int B, int L, int Cin, int Lout, int k,
                               int pad_left, void* stream);
int gn_conv1d_cout1_wgrad_bf16(const void* x, const float* dy, float* dw, float* db, int B, int L, int Cin, int Lout, int k,
                               int pad_left, void* stream);
/* UpSampling1D on bf16 activations (C % 8 == 0): y (B, L*size, C) from x (B, L, C); bwd sums the `size` copies */
int gn_upsample1d_fwd_bf16(const void* x, void* y, int B, int L, int C, int size, void* stream);
int gn_upsample1d_bwd_bf16(const void* dy, void* dx, int B, int L, int C, int size, void* stream);
int gn_dense_small_fwd_bf16(const void* x, const float* w, const float* bias, float* y, int M, int K, int N, int act,
                            float act_param, void* stream);
int gn_dense_small_dgrad_bf16(const float* dy, const float* w, const void* x_in, void* dx, float* dx_colsum,
                              int colsum_channels, int M, int K, int N, int in_act, float in_act_param, void* stream);
int gn_dense_small_wgrad_bf16(const void* x, const float* dy, float* dw, float* db, int M, int K, int N, void* stream);
int gn_act_bwd_bf16(const void* dy, const void* y, void* dx, long long n, int act, float param, void* stream);

/* Conv2D(5x5, strides (2,1), 'same') over an (L,2,C) image (bbhMahoGANy.py:439,447) is a Conv1D with
 * Cin'=2Cin, Cout'=2Cout: w1 (kh, 2Cin, 2Cout) [kh,(wi,ci),(wo,co)] = w2 (kh,kw,Cin,Cout) [kh, wi-wo+pw, ci, co];
 * pw = left width pad (2 for kw=5). pack: w2->w1, b (Cout)->b1 (2Cout); unpack: dw1->dw2, db1->db (overwrite). */
int gn_conv2d_w2_pack_f32(const float* w2, const float* b, float* w1, float* b1, int kh, int kw, int Cin,
                          int Cout, int pw, void* stream);
int gn_conv2d_w2_unpack_f32(const float* dw1, const float* db1, float* dw2, float* db, int kh, int kw, int Cin,
                            int Cout, int pw, void* stream);

/* Dense: y (M,N) = act(x (M,K) @ w (K,N) + bias) */
int gn_dense_fwd_f32(const float* x, const float* w, const float* bias, float* y, int M, int K, int N, int act,
                     float act_param, void* stream);
int gn_dense_dgrad_f32(const float* dy, const float* w, float* dx, int M, int K, int N, void* stream);
int gn_dense_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int M, int K, int N, void* stream);

/* BatchNormalization(axis=-1) over `rows` x C (Keras 2.2.4: eps 1e-3, biased batch variance for the
 * normalisation, moving_var fed var*n/(n-(1+eps)), bbhMahoGANy.py:235,251).
 * stats: f32 workspace (2*C): on return [0:C]=batch mean, [C:2C]=1/sqrt(var+eps) (kept for backward).
 * stats_only!=0 stops after the per-channel sums so a data-parallel caller can all-reduce them:
 *   step 0: gn_bn_sums(x) -> sums (2C) = (sum x, sum x^2 about `shift`)  ... see gn_bn_* below. */
int gn_bn_stats_f32(const float* x, long long rows, int C, double* sums /* (2C): sum(x), sum((x-shift)^2) */,
                    const float* shift /* (C) or NULL */, void* stream);
/* phase 0: stats[0:C] = sum_x / n.  phase 1: stats[C:2C] = 1/sqrt(sum_sq/n + eps) and, when moving_mean / moving_var
 * are given, their update.  biased == NULL: plain exponential average moving = moving*m + batch*(1-m).
 * biased != NULL (f32 (2C), zero-initialised by the caller, one per layer): the zero-debiased average Keras 2.2.4
 * gets from TF 1.12 (K.moving_average_update -> assign_moving_average(zero_debias=True)):
 * biased = biased*m + batch*(1-m); moving = biased * debias with debias = 1/(1 - m^step), step counted by the caller. */
int gn_bn_finalize_f32(const double* sum_x, const double* sum_sq, double n_total, int C, float eps, float momentum,
                       float* stats, float* moving_mean, float* moving_var, int phase, float* biased, double debias,
                       void* stream);
int gn_bn_apply_f32(const float* x, const float* mean, const float* invstd_or_var, const float* gamma,
                    const float* beta, float* y, long long rows, int C, float eps, int use_var, void* stream);
/* backward sums: sums (2C) double = (sum dy, sum dy*xhat) with xhat=(x-mean)*invstd */
int gn_bn_bwd_sums_f32(const float* x, const float* dy, const float* stats, long long rows, int C, double* sums,
                       void* stream);
/* dx = gamma*invstd*(dy - sum_dy/n - xhat*sum_dyxhat/n); dgamma=sum_dyxhat, dbeta=sum_dy (overwrite) */
int gn_bn_bwd_apply_f32(const float* x, const float* dy, const float* stats, const float* gamma, const double* sums,
                        double n_total, float* dx, float* dgamma, float* dbeta, long long rows, int C,
                        void* stream);

/* bf16 throughput mode of BatchNormalization -> activation -> dropout chains (generator hidden layers,
 * bbhMahoGANy.py:235-289) as streaming passes over bf16 activations, fp32 arithmetic, double statistics; C % 8 == 0.
 *   gn_bn_stats_bf16      : sums (2C) double OVERWRITTEN = (sum x, sum x^2) per channel (one pass; data-parallel callers
 *                           all-reduce it, then gn_bn_finalize_f32 with sum_sq - sum^2/n gives mean / invstd / moving stats)
 *   gn_chain_fwd_bf16     : y = noise(act(gamma * (x - mean) * invstd + beta)); scale = invstd, or the moving variance
 *                           when use_var (inference); mean == NULL skips the normalisation (plain activation / dropout);
 *                           noise = -1 | GN_NOISE_DROPOUT | GN_NOISE_GDROPOUT with the mask from r (f32, fed) or, when r is
 *                           NULL, from the Philox stream of gn_noise_draw_f32(seed, offset) -- never materialised
 *   gn_chain_bwd_sums_bf16: sums (2C) double OVERWRITTEN = (sum g, sum g*xhat), g = dy * noise' * act'(a), a recomputed
 *   gn_chain_bwd_bf16     : dx = gamma*invstd*(g - sum_g/n - xhat*sum_gxhat/n), dgamma = sum_gxhat, dbeta = sum_g
 *                           (mean == NULL: dx = g) */
int gn_bn_stats_bf16(const void* x, long long rows, int C, double* sums, void* stream);
int gn_chain_fwd_bf16(const void* x, void* y, const float* mean, const float* scale, const float* gamma, const float* beta,
                      int use_var, float eps, int act, float act_param, int noise, float rate, const float* r,
                      uint64_t seed, uint64_t offset, long long rows, int C, void* stream);
int gn_chain_bwd_sums_bf16(const void* x, const void* dy, const float* mean, const float* invstd, const float* gamma,
                           const float* beta, int act, float act_param, int noise, float rate, const float* r,
                           uint64_t seed, uint64_t offset, long long rows, int C, double* sums, void* stream);
int gn_chain_bwd_bf16(const void* x, const void* dy, void* dx, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, const double* sums, double n_total, int act, float act_param, int noise,
                      float rate, const float* r, uint64_t seed, uint64_t offset, float* dgamma, float* dbeta,
                      long long rows, int C, void* stream);

/* The same chains over FLOAT32 activations (float32 / split-operand modes): libm-accurate activations, per-thread
 * statistics in double; gn_bn_sums_f32 is the one-pass (sum x, sum x^2) form of gn_bn_stats_bf16. */
int gn_bn_sums_f32(const float* x, long long rows, int C, double* sums, void* stream);
int gn_chain_fwd_f32(const float* x, float* y, const float* mean, const float* scale, const float* gamma, const float* beta,
                     int use_var, float eps, int act, float act_param, int noise, float rate, const float* r,
                     uint64_t seed, uint64_t offset, long long rows, int C, void* stream);
int gn_chain_bwd_sums_f32(const float* x, const float* dy, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, int act, float act_param, int noise, float rate, const float* r,
                          uint64_t seed, uint64_t offset, long long rows, int C, double* sums, void* stream);
int gn_chain_bwd_f32(const float* x, const float* dy, float* dx, const float* mean, const float* invstd, const float* gamma,
                     const float* beta, const double* sums, double n_total, int act, float act_param, int noise,
                     float rate, const float* r, uint64_t seed, uint64_t offset, float* dgamma, float* dbeta,
                     long long rows, int C, void* stream);
/* the same two apply passes with a side output: y_amax / dx_amax (device scalar, OVERWRITTEN) = max |result|, the scale
 * source of the consumer's gn_split_f32_f16x2(have_amax = 1) */
int gn_chain_fwd_amax_f32(const float* x, float* y, const float* mean, const float* scale, const float* gamma, const float* beta,
                          int use_var, float eps, int act, float act_param, int noise, float rate, const float* r,
                          uint64_t seed, uint64_t offset, long long rows, int C, float* y_amax, void* stream);
/* forward apply pass that ALSO writes its result as the scaled fp16 pair the next convolution consumes: y_planes fp16
 * (2, rows, C), scaled for the a-priori bound `bound` >= max |y| (bounded activations: tanh, sigmoid, ReLU(max_value),
 * times the dropout factor 1 / (1 - rate)); y_amax[0] = bound on return.  Saves the separate split pass over y. */
int gn_chain_fwd_planes_f32(const float* x, float* y, const float* mean, const float* scale, const float* gamma,
                            const float* beta, int use_var, float eps, int act, float act_param, int noise, float rate,
                            const float* r, uint64_t seed, uint64_t offset, long long rows, int C, void* y_planes,
                            float* y_amax, float bound, void* stream);
int gn_chain_bwd_amax_f32(const float* x, const float* dy, float* dx, const float* mean, const float* invstd,
                          const float* gamma, const float* beta, const double* sums, double n_total, int act, float act_param,
                          int noise, float rate, const float* r, uint64_t seed, uint64_t offset, float* dgamma, float* dbeta,
                          long long rows, int C, float* dx_amax, void* stream);

/* elementwise */
int gn_act_fwd_f32(const float* x, float* y, long long n, int act, float param, void* stream);
int gn_act_bwd_f32(const float* dy, const float* y, float* dx, long long n, int act, float param, void* stream);
int gn_noise_fwd_f32(const float* x, const float* r, float* y, long long n, int kind, float rate, void* stream);
int gn_noise_bwd_f32(const float* dy, const float* r, float* dx, long long n, int kind, float rate, void* stream);
/* fill r for the dropout family with Philox4x32-10 (kind DROPOUT: keep mask with P(keep)=1-rate; else N(0,1)) */
int gn_noise_draw_f32(float* r, long long n, int kind, float rate, uint64_t seed, uint64_t offset, void* stream);
int gn_uniform_f32(float* r, long long n, float lo, float hi, uint64_t seed, uint64_t offset, void* stream);
int gn_normal_f32(float* r, long long n, float mean, float std, uint64_t seed, uint64_t offset, void* stream);
int gn_upsample1d_fwd_f32(const float* x, float* y, int B, int L, int C, int size, void* stream);
int gn_upsample1d_bwd_f32(const float* dy, float* dx, int B, int L, int C, int size, void* stream);
int gn_maxpool1d_fwd_f32(const float* x, float* y, int B, int L, int C, int pool, void* stream);
int gn_maxpool1d_bwd_f32(const float* x, const float* y, const float* dy, float* dx, int B, int L, int C, int pool,
                         void* stream);
/* Layers of the 2_model_version networks.
 * gn_gap_{fwd,bwd}_f32: GlobalAveragePooling1D / 2D (no_weight_code/subtract_model.py:330,371): y (B,C) = mean over the
 *   L positions of x (B,L,C); dx = dy / L broadcast.
 * gn_transpose_f32: y (B,C,R) = x (B,R,C); BatchNormalization(axis=1) (no_weight_code/subtract_model.py:264-292,344-357)
 *   moves the normalised axis innermost, runs the channels-last kernels and moves it back.
 * gn_reg_terms_f32: keras.regularizers.l1 / l2 (weight_version/subtract_model.py:217): *loss (double, ACCUMULATED; may be
 *   NULL) += l1 sum|x| + l2 sum x^2 and, when g != NULL, g += l1 sign(x) + 2 l2 x. */
int gn_gap_fwd_f32(const float* x, float* y, int B, int L, int C, void* stream);
int gn_gap_bwd_f32(const float* dy, float* dx, int B, int L, int C, void* stream);
int gn_transpose_f32(const float* x, float* y, int B, int R, int C, void* stream);
int gn_reg_terms_f32(const float* x, float* g, long long n, float l1, float l2, double* loss, void* stream);
/* a += b (gradient accumulation where two branches share an input, bbhMahoGANy.py:362,382) */
int gn_axpy_f32(float* a, const float* b, float alpha, long long n, void* stream);
/* Evaluation stage of the GAN loop (bbhMahoGANy.py:1311-1343 -> make_contour_plot :787-791, overlap_tests :853-870).
 * gn_kde2d_pdf_f32: scipy.stats.gaussian_kde(dataset).pdf(positions) for a two-dimensional dataset:
 *   pdf[j] = inv_norm * sum_i exp(-1/2 (x_i - p_j)^T A (x_i - p_j)),  A = [[a11,a12],[a12,a22]] = inverse of the
 *   bandwidth-scaled data covariance (Scott factor n^(-1/6)), inv_norm = 1 / (n sqrt(det(2 pi covariance))).
 * data_xy (n,2) f32 and pos_xy (m,2) f32 interleaved (x,y) pairs (the host centres both on the data mean);
 * pdf (m) f32.
 * gn_overlap_sums_f32: out3 = {sum a b, sum a^2, sum b^2} (double), beta = out3[0] / sqrt(out3[1] out3[2]) (:868-870). */
int gn_kde2d_pdf_f32(const float* data_xy, int n, const float* pos_xy, int m, double a11, double a12, double a22,
                     double inv_norm, float* pdf, void* stream);
int gn_overlap_sums_f32(const float* a, const float* b, long long n, double* out3, void* stream);
/* Percentile curves of plot_waveform_est (bbhMahoGANy.py:913-921): out (npct, L) f32, out[q, l] = np.percentile(x[:, l],
 * pcts[q]) (default linear interpolation) of the n generated waveforms x (n, L) f32; n <= 32768, npct <= 64. */
int gn_percentiles_f32(const float* x, int n, int L, const float* pcts, int npct, float* out, void* stream);
/* gather rows: out[i,:] = src[idx[i],:]  (template batch assembly, bbhMahoGANy.py:1156-1158,1244) */
int gn_gather_rows_f32(const float* src, const int* idx, float* out, int n, long long row_len, void* stream);

/* Conv2DTranspose((1,kw), strides (1,1), 'valid') of the 2_model_version generators
 * (2_model_version/weight_version/no_mode_collapse_network.py:79-90) runs on the Conv1D entry points with
 * pad_left = kw-1, Lout = L+kw-1 and the kernel W1[t,ci,co] = K[0,kw-1-t,co,ci]; this converts between the two
 * layouts (weights: A=Cin,B=Cout from Keras (kw,Cout,Cin); gradients: A=Cout,B=Cin back): out[t,a,b] = in[k-1-t,b,a]. */
int gn_flip_transpose_f32(const float* in, float* out, int k, int A, int B, void* stream);

/* waveform ingest, train_on_wvf_version/load_txtwfs.py:47-50,66-69: y[b, (j + offsets[b]) mod N] = x[b, j] / max_j x[b, j]
 * (np.max then np.roll; offsets may be negative or NULL = no shift).  The FFT resampling step before it
 * (scipy.signal.resample(data, 512)) is a fixed linear map for a given input length and runs as gn_dense_fwd_f32
 * against the precomputed (Nx, 512) Dirichlet-kernel matrix. */
int gn_maxnorm_roll_f32(const float* x, const int* offsets, float* y, int B, int N, void* stream);

/* MyLayer, bbhMahoGANy.py:164-188: y (B,L,2) = stack([x, const - x], axis=2); bwd dx = dy[...,0]-dy[...,1] */
int gn_stack_residual_fwd_f32(const float* x, const float* cst, float* y, int B, int L, void* stream);
int gn_stack_residual_bwd_f32(const float* dy, float* dx, int B, int L, void* stream);
/* real images of the GAN loop, bbhMahoGANy.py:1276-1284: y (n,2) = concatenate((signal, noise), axis=2) */
int gn_stack_pair_f32(const float* a, const float* b, float* y, long long n, void* stream);
/* MyLayer, tests/burstMahoGANy.py:100-125: out (2) = [mean(c-x), mean((c-x)^2)] over the whole batch;
 * sums (2) double workspace = un-normalised sums (for data-parallel all-reduce) */
int gn_residual_moments_fwd_f32(const float* x, const float* cst, double* sums, int B, int L, void* stream);
int gn_residual_moments_bwd_f32(const float* x, const float* cst, const float* dout /* (2) */, float* dx, int B,
                                int L, double n_total, void* stream);

/* losses: pred (B,D), target (B,D) or (B) broadcast... all as (B,D) f32.
 * out (2) f32: [0] += sum_b mean_d loss, [1] += sum_b metric hits  (caller zeroes, divides by global B)
 * dpred (B,D) = d(mean_b loss)/dpred using inv_batch = 1/global batch. pred_is_vec: pred is (D) broadcast over B
 * (burst MyLayer output), dpred then (D) accumulated. */
int gn_loss_fwd_bwd_f32(const float* pred, const float* target, float* out, float* dpred, int B, int D, int kind,
                        float param, float inv_batch, int pred_is_vec, int metric_kind, void* stream);

/* Keras optimizers over a flat parameter arena (one launch per model):
 * Adam (Keras 2.2.4): m=b1 m+(1-b1)g; v=b2 v+(1-b2)g^2; p-=lr_t*m/(sqrt(v)+eps), lr_t computed by the caller */
int gn_adam_step_f32(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1, float beta2,
                     float eps, float grad_scale, void* stream);
int gn_sgd_step_f32(float* p, const float* g, long long n, float lr, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENNET_B200_H */
