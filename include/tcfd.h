/* tcfd.h -- C ABI of the B200-native torch-cfd spectral hot path (libtcfd.so).
 *
 * The reference (scaomath/torch-cfd) is pure Python and has no FFI seam for this path: the
 * seam is its nn.Module API.  Each entry point below therefore names the reference *method* it
 * stands behind; the Python shims in torch-cfd_b200/ keep those methods' names and signatures and
 * forward to these functions through ctypes (see INTEGRATION.md for the binding a maintainer
 * would add upstream).
 *
 * Conventions
 *  - Plain C: pointers + sizes only, no torch types.  Device pointers are raw CUDA addresses
 *    (torch: tensor.data_ptr()); the library never frees or reallocates caller memory.
 *  - Complex arrays are interleaved (re, im) pairs of float (prec 32) or double (prec 64), in
 *    the reference's layout: spectrum (B, n, n/2+1) row-major, "backward" normalisation.
 *  - All kernels are enqueued on the caller's stream (cudaStream_t passed as void*); calls are
 *    asynchronous with respect to the host and never synchronise.
 *  - Every function returns 0 on success or a negative tcfd error code; tcfd_last_error()
 *    returns a thread-local human-readable message.  Nothing throws or exits.
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TCFD_H_
#define TCFD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCFD_OK 0
#define TCFD_ERR_INVALID (-1) /* bad argument / unsupported size */
#define TCFD_ERR_CUDA (-2)    /* a CUDA runtime call failed */
#define TCFD_ERR_NOMEM (-3)

const char* tcfd_last_error(void);
/* library version / build string, e.g. "tcfd 0.1 sm_100a" */
const char* tcfd_version(void);

/* ------------------------------------------------------------------------------------------
 * Hot path A: pseudo-spectral 2-D vorticity Navier-Stokes, RK4 (Carpenter-Kennedy) + CN.
 * Replaces: NavierStokes2DSpectral (torch_cfd/equations.py:361-463) driven by
 *           RK4CrankNicolsonStepper.forward (torch_cfd/equations.py:328-358).
 * ---------------------------------------------------------------------------------------- */
typedef struct tcfd_ns2d tcfd_ns2d_t;

typedef struct {
  int n;         /* grid is n x n, n a power of two in [32, 2048] (fp64: up to 1024) */
  int prec;      /* 32 or 64: precision of every table and of the state */
  int max_batch; /* workspace is sized for this many samples */
  /* HOST pointers, copied at creation.  They are the reference's own buffers
   * (NavierStokes2DSpectral._initialize, torch_cfd/equations.py:394-403), so that rounding of the
   * tables is the reference's, not ours: */
  const void* kappa_x;     /* [n]        imag(2j*pi*kx[:,0])   (torch_cfd/equations.py:417)        */
  const void* kappa_y;     /* [n/2+1]    imag(2j*pi*ky[0,:])                                        */
  const void* neg_inv_lap; /* [n][n/2+1] -1/laplace' with laplace'[0,0]=1 (torch_cfd/spectral.py:41-46,113) */
  const void* linear_term; /* [n][n/2+1] viscosity*laplace - drag       (torch_cfd/equations.py:401)*/
  const void* filter;      /* [n][n/2+1] 2/3 brick-wall mask, or NULL when smooth=False
                              (torch_cfd/spectral.py:78-84, torch_cfd/equations.py:424-425)         */
  const void* f_hat;       /* [n][n/2+1] complex forcing spectrum added to F, or NULL
                              (torch_cfd/equations.py:429-437)                                      */
} tcfd_ns2d_desc_t;

int tcfd_ns2d_create(tcfd_ns2d_t** out, const tcfd_ns2d_desc_t* desc);
int tcfd_ns2d_destroy(tcfd_ns2d_t* h);
/* replace the forcing spectrum (HOST pointer, NULL = no forcing) */
int tcfd_ns2d_set_forcing(tcfd_ns2d_t* h, const void* f_hat);
/* bytes of device workspace owned by the handle */
size_t tcfd_ns2d_workspace_bytes(const tcfd_ns2d_t* h);
/* 0, or TCFD_ERR_CUDA when a kernel of an earlier (asynchronous) call on this handle reported a
 * failure through the handle's host-visible error word; meaningful after the stream was synchronised.
 * Every tcfd_ns2d_step call performs this check on entry. */
int tcfd_ns2d_check(const tcfd_ns2d_t* h);
/* launch schedule of tcfd_ns2d_step on this handle: 1 = ONE launch per call -- the persistent dataflow kernel
 * (n >= 256; environment TCFD_FLOW=0 at creation selects the other) or the shared-memory-resident kernel
 * (n <= 64; TCFD_SMALL=0 selects the other) --, 0 = two launches per substage */
int tcfd_ns2d_schedule(const tcfd_ns2d_t* h);
/* number of kernel launches the last tcfd_ns2d_* compute call enqueued */
int tcfd_ns2d_last_launch_count(const tcfd_ns2d_t* h);

/* NavierStokes2DSpectral.forward(vort_hat, dt, steps) with an RK4CrankNicolsonStepper
 * (torch_cfd/equations.py:452-463, :328-358).
 *   w_in   [batch][n][n/2+1] complex, device, read-only
 *   w_out  same shape, device, must not alias w_in
 *   dwdt   same shape or NULL: (w_out - w_in) * inv_total_dt     (equations.py:462)
 *   nstages, beta[nstages], gdt[nstages] = gammas[k]*dt, mu[nstages] = 0.5*dt*(alphas[k+1]-alphas[k]):
 *          the stepper's coefficients, evaluated by the caller in the stepper's dtype exactly as
 *          equations.py:355-357 does, passed as doubles.
 */
int tcfd_ns2d_step(tcfd_ns2d_t* h, const void* w_in, void* w_out, void* dwdt, int batch, int steps,
                   int nstages, const double* beta, const double* gdt, const double* mu,
                   double inv_total_dt, void* stream);

/* Measurement twin of tcfd_ns2d_step (no reference counterpart): brackets EVERY kernel launch of
 * the step with CUDA events on `stream`, synchronises the stream, and returns the summed device
 * time ms[4] and launch count count[4] per kernel kind (0 = rows-inverse prologue, 1 = rows
 * forward+inverse -- or, under the dataflow schedule, the single persistent launch of the call --
 * 2 = rows forward epilogue, 3 = cols).  Used by bench.py for the roofline. */
int tcfd_ns2d_step_timed(tcfd_ns2d_t* h, const void* w_in, void* w_out, int batch, int steps, int nstages,
                         const double* beta, const double* gdt, const double* mu, void* stream, float* ms,
                         int* count);

/* NavierStokes2DSpectral.explicit_terms(vort_hat)  (torch_cfd/equations.py:413-441) */
int tcfd_ns2d_explicit_terms(tcfd_ns2d_t* h, const void* w_in, void* f_out, int batch, void* stream);

/* NavierStokes2DSpectral.residual(vhat, vt_hat) = vt_hat - F(vhat) - L vhat
 * (torch_cfd/equations.py:405-411) */
int tcfd_ns2d_residual(tcfd_ns2d_t* h, const void* w_in, const void* wt_in, void* r_out, int batch,
                       void* stream);

/* Recording step of get_trajectory_imex (fno/data_gen/solvers.py:245-256): writes slot `it` of the
 * device snapshot buffers [batch][n_t][n][n/2+1] (complex64 when out_prec = 32, complex128 when 64):
 *   snap_w <- w, snap_psi <- -1/laplace' * w (vorticity_to_velocity, torch_cfd/spectral.py:113),
 *   snap_dwdt <- dwdt, snap_res <- res.  Any snapshot pointer may be NULL. */
int tcfd_ns2d_record(tcfd_ns2d_t* h, const void* w, const void* dwdt, const void* res, void* snap_w,
                     void* snap_psi, void* snap_dwdt, void* snap_res, int batch, int n_t, int it, int out_prec,
                     void* stream);

/* Same as tcfd_ns2d_step but with HOST buffers (pinned memory recommended): uploads w_in,
 * steps, downloads w_out (and dwdt if not NULL) on `stream`, batch-chunked so copies overlap
 * compute.  This is the end-to-end call a host-resident caller makes. */
int tcfd_ns2d_step_host(tcfd_ns2d_t* h, const void* w_in_host, void* w_out_host, void* dwdt_host,
                        int batch, int steps, int nstages, const double* beta, const double* gdt,
                        const double* mu, double inv_total_dt, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hot path B: FNO3d / SFNO spectral convolution  y = irfftn( W (.) rfftn(x) )  on the four retained
 * corner blocks, fp32.
 * Replaces: SpectralConv3d.forward (fno/fno3d.py:86-116), SpectralConv.forward (fno/base.py:229-237)
 *           with SpectralConvS.spectral_conv (fno/sfno.py:364-391), SpectralConvT.forward
 *           (fno/sfno.py:433-457, postprocess = Identity), and their autograd backward.
 *   x        [batch][Ci][X][Y][T_in]  float, device           (time innermost, as the reference)
 *   y        [batch][Co][X][Y][T_out] float, device
 *   w[4]     four device pointers, each [Ci][Co][mx][my][mt] complex64: weights1..4 of SpectralConv3d
 *            = view_as_complex(weight[0..3]) of SpectralConvS -- corner order (lo x, lo y),
 *            (hi x, lo y), (lo x, hi y), (hi x, hi y)
 *   bias[4]  NULL or four device pointers [mx][my][mt] complex64; delta * bias is added to every
 *            (batch, out-channel) entry of the corner (fno/sfno.py:386-388)
 * ---------------------------------------------------------------------------------------- */
typedef struct tcfd_sconv3d tcfd_sconv3d_t;

typedef struct {
  int X, Y;      /* spatial grid: powers of two in [32, 512] */
  int T_in;      /* time samples of x */
  int t_pad;     /* zeros prepended in time before the transform (SpectralConvT temporal_padding), else 0 */
  int T_out;     /* time samples of y: irfftn(s=(X, Y, T_out + t_pad)) keeping the last T_out */
  int Ci, Co;
  int mx, my, mt; /* retained modes: mx <= X/2, my <= Y/2, mt <= (T_in + t_pad)/2 + 1 */
  int norm;      /* 0 "backward" (torch default), 1 "ortho", 2 "forward" */
  int max_batch;
} tcfd_sconv3d_desc_t;

int tcfd_sconv3d_create(tcfd_sconv3d_t** out, const tcfd_sconv3d_desc_t* desc);
int tcfd_sconv3d_destroy(tcfd_sconv3d_t* h);
size_t tcfd_sconv3d_workspace_bytes(const tcfd_sconv3d_t* h);
/* complex64 elements of the truncated input spectrum saved for backward: batch * Ci * 4 mx my mt */
size_t tcfd_sconv3d_xhat_elems(const tcfd_sconv3d_t* h, int batch);
int tcfd_sconv3d_last_launch_count(const tcfd_sconv3d_t* h);

/* forward; xhat_save (device, tcfd_sconv3d_xhat_elems complex64) may be NULL when no backward follows */
int tcfd_sconv3d_forward(tcfd_sconv3d_t* h, const void* x, const void* const* w, const void* const* bias,
                         float delta, void* y, void* xhat_save, int batch, void* stream);

/* backward of the above for a cotangent grad_y [batch][Co][X][Y][T_out]:
 *   grad_x   [batch][Ci][X][Y][T_in] or NULL
 *   grad_w   NULL or four pointers shaped like w        (torch convention: sum_b conj(x_hat) g_hat)
 *   grad_bias NULL or four pointers shaped like bias    (delta * sum over batch and out-channels) */
int tcfd_sconv3d_backward(tcfd_sconv3d_t* h, const void* grad_y, const void* xhat, const void* const* w,
                          void* grad_x, void* const* grad_w, void* const* grad_bias, float delta, int batch,
                          void* stream);

/* The two halves of the layer as separate calls, for callers that put something between them in spectral
 * space: SpectralConvT's `postprocess` (HelmholtzProjection, fno/sfno.py:116-193, :452) or a change of mesh
 * (SpectralConv.forward with out_mesh_size, fno/base.py:229-237 -- the synthesis then runs on the handle of the
 * output geometry).  yhat = the truncated output spectrum [batch][Co][2mx][2my][mt] complex64 (kx rows: the mx
 * lowest then the mx highest frequencies, likewise ky), in the layer's normalisation:
 *   forward  == synthesis(analysis(x));  backward == analysis_backward(synthesis_backward(grad_y)).
 * grad_yhat follows torch's convention for gradients of complex tensors. */
size_t tcfd_sconv3d_yhat_elems(const tcfd_sconv3d_t* h, int batch);
int tcfd_sconv3d_analysis(tcfd_sconv3d_t* h, const void* x, const void* const* w, const void* const* bias, float delta,
                          void* yhat, void* xhat_save, int batch, void* stream);
int tcfd_sconv3d_synthesis(tcfd_sconv3d_t* h, const void* yhat, void* y, int batch, void* stream);
int tcfd_sconv3d_synthesis_backward(tcfd_sconv3d_t* h, const void* grad_y, void* grad_yhat, int batch, void* stream);
int tcfd_sconv3d_analysis_backward(tcfd_sconv3d_t* h, const void* grad_yhat, const void* xhat, const void* const* w,
                                   void* grad_x, void* const* grad_w, void* const* grad_bias, float delta, int batch,
                                   void* stream);

/* ------------------------------------------------------------------------------------------
 * FNO3d layer glue: the pointwise channel mixes around the spectral convolution,
 * inference only, fp32, tensors (batch, C, X, Y, T) contiguous with npts = X*Y*T points per channel plane.
 * Weights are in torch's Conv3d layout [Co][Ci] (kernel size 1), biases [Co] or NULL: DEVICE pointers for
 * tcfd_fno_pointwise_linear and tcfd_fno_project, HOST pointers for tcfd_fno_layer_glue (its three C x C
 * matrices travel as a kernel parameter, so that the products read them from the constant bank).
 * Replaces, in FNO3d.forward (fno/fno3d.py:205-236):
 *   tcfd_fno_pointwise_linear   x = self.p(x)                                   (:214, Conv3d 1x1x1)
 *   tcfd_fno_layer_glue         x = nonlinear( mlp(conv(x)) + w(x) )            (:223-230; mlp = MLP :119-130:
 *                               mlp2(gelu(mlp1(.))), nonlinear = GELU (act 1) or Identity (act 0))
 *   tcfd_fno_project            x = self.q(x)                                   (:235; MLP with Co = 1 and
 *                               hidden width M, GELU between when act = 1)
 * ---------------------------------------------------------------------------------------- */
int tcfd_fno_pointwise_linear(const float* x, float* y, const float* w, const float* bias, int batch, int Ci, int Co,
                              size_t npts, void* stream);
int tcfd_fno_layer_glue(const float* conv_out, const float* x, float* y, const float* w1, const float* b1,
                        const float* w2, const float* b2, const float* ww, const float* bw, int act, int batch, int C,
                        size_t npts, void* stream);
/* which kernel the last tcfd_fno_layer_glue call of this process launched: 1 = tensor cores (tcgen05 3xTF32 products,
 * accumulators in TMEM: even C <= 32 on a CUDA device), 0 = CUDA cores (odd C, TCFD_GLUE_TC=0, host-emulation build) */
int tcfd_fno_layer_glue_path(void);
int tcfd_fno_project(const float* x, float* y, const float* w1, const float* b1, const float* w2, const float* b2,
                     int act, int batch, int C, int M, size_t npts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone batched 2-D real transforms and the bilinear resampling of the data-generation scripts
 * (post-processing of recorded trajectories on the device; also the forcing spectra and
 * initial-condition generators).  Layouts as above: spectrum [count][n][n/2+1] complex, field [count][n][n]
 * real, precision of the handle, "backward" normalisation.
 *   tcfd_fft2_irfft2         torch.fft.irfft2(value)    (fno/data_gen/data_gen_Kolmogorov2d.py:178-179,
 *                            data_gen_McWilliams2d.py:157): C2R semantics included (imaginary parts of the
 *                            ky = 0, n/2 bins are dropped after the kx transform)
 *   tcfd_fft2_rfft2          torch.fft.rfft2(field)     (torch_cfd/equations.py:432-436)
 *   tcfd_resample_bilinear   F.interpolate(value, size=(n_out, n_out), mode="bilinear")
 *                            (fno/data_gen/data_gen_Kolmogorov2d.py:186, align_corners=False) with the cast
 *                            `.to(dtype)` of :179 folded in (prec_in -> prec_out)
 * ---------------------------------------------------------------------------------------- */
typedef struct tcfd_fft2 tcfd_fft2_t;
int tcfd_fft2_create(tcfd_fft2_t** out, int n, int prec);
int tcfd_fft2_destroy(tcfd_fft2_t* h);
int tcfd_fft2_last_launch_count(const tcfd_fft2_t* h);
int tcfd_fft2_irfft2(tcfd_fft2_t* h, const void* in_hat, void* out, int count, void* stream);
int tcfd_fft2_rfft2(tcfd_fft2_t* h, const void* in, void* out_hat, int count, void* stream);
int tcfd_resample_bilinear(const void* in, void* out, int prec_in, int prec_out, int count, int n_in, int n_out,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TCFD_H_ */
