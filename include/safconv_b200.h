/*
 * safconv_b200.h -- C ABI of libsafconv_b200.so
 *
 * B200 (sm_100a) implementation of the Spatial_Audio_Framework multichannel
 * convolution path.  The first block of declarations is a DROP-IN for the
 * reference header
 *
 *   /root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.h
 *
 * (same symbol names, same signatures, same data layouts, same ownership
 * rules); a program that links libsafconv_b200.so instead of the reference's
 * saf_utility_matrixConv.o needs no source change.  The second block is an
 * additive extension surface (device-pointer apply, batched frames, sharding,
 * error query, timing) that the reference does not have.
 *
 * There is NO CPU fallback: if no CUDA device is usable, `*_create` stores NULL
 * in *phMC, records an error (see safconv_last_error_string) and returns;
 * `*_apply` on a NULL handle is a no-op, exactly like the reference's callers
 * already assume (examples/src/matrixconv/matrixconv.c:141-145).
 *
 * All pointers are plain host pointers to contiguous float32 unless a function
 * name ends in `_device`.
 */
#ifndef SAFCONV_B200_H_INCLUDED
#define SAFCONV_B200_H_INCLUDED

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ========================================================================== */
/*            Drop-in block (reference: saf_utility_matrixConv.h)             */
/* ========================================================================== */

/**
 * Matrix convolver: nCHout x nCHin FIR matrix, block-by-block.
 * Replaces saf_matrixConv_create, saf_utility_matrixConv.h:55-62
 * (implementation saf_utility_matrixConv.c:49-130).
 *
 * @param phMC        (&) handle; overwritten (NULL on failure)
 * @param hopSize     block length in samples, any positive value for saf_matrixConv / saf_multiConv (up to 8192 on the
 *                    partitioned engine; larger blocks on the big-FFT engine of csrc/safconv_np.c, same linear convolution --
 *                    only an ODD numOvrlpAddBlocks * hopSize above 8192 is rejected); saf_TVConv: 1..8192, larger values are
 *                    rejected with an error string (the reference's tvconv host clamps its frames to 8192)
 * @param H           time-domain filters, FLAT nCHout x nCHin x length_h; only
 *                    read during this call (caller may free it afterwards)
 * @param usePartFLAG 0/1 as in the reference.  Both modes produce the same causal
 *                    linear convolution; this library serves both with its
 *                    uniformly-partitioned engine (outputs equal to rounding).
 */
void saf_matrixConv_create(void** const phMC, int hopSize, float* H, int length_h,
                           int nCHin, int nCHout, int usePartFLAG);

/** Replaces saf_matrixConv_destroy, saf_utility_matrixConv.h:69-70 (.c:132-161).
 *  Safe on *phMC == NULL.  Additionally sets *phMC = NULL (harmless superset). */
void saf_matrixConv_destroy(void** const phMC);

/**
 * Replaces saf_matrixConv_apply, saf_utility_matrixConv.h:82-86 (.c:163-236).
 * Synchronous: outputSigs (FLAT nCHout x hopSize) is complete on return.
 * inputSigs is FLAT nCHin x hopSize.
 */
void saf_matrixConv_apply(void* const hMC, float* inputSigs, float* outputSigs);

/**
 * Multi-channel (diagonal) convolver: nCH independent FIRs.
 * Replaces saf_multiConv_create, saf_utility_matrixConv.h:109-115 (.c:257-328).
 * H is FLAT nCH x length_h.
 */
void saf_multiConv_create(void** const phMC, int hopSize, float* H, int length_h,
                          int nCH, int usePartFLAG);

/** Replaces saf_multiConv_destroy, saf_utility_matrixConv.h:122-123 (.c:330-355). */
void saf_multiConv_destroy(void** const phMC);

/** Replaces saf_multiConv_apply, saf_utility_matrixConv.h:132-136 (.c:357-414).
 *  inputSigs / outputSigs are FLAT nCH x hopSize. */
void saf_multiConv_apply(void* const hMC, float* inputSigs, float* outputSigs);

/**
 * Time-varying convolver (1 input -> nCHout, nIRs selectable impulse responses,
 * 3-way cross-fade on IR changes).
 * Replaces saf_TVConv_create / _destroy / _apply,
 * saf_utility_matrixConv.h:157-190 (.c:423-620).
 * H: nIRs pointers, each to FLAT nCHout x length_h.
 */
void saf_TVConv_create(void** const phTVC, int hopSize, float** H, int length_h,
                       int nIRs, int nCHout, int initIdx);
void saf_TVConv_destroy(void** const phTVC);
void saf_TVConv_apply(void* const hTVC, float* inputSigs, float* outputSigs, int irIdx);

/* ========================================================================== */
/*                 Extension block (not present in the reference)             */
/* ========================================================================== */

/**
 * usePartFLAG = 0 with the reference's TRUE non-partitioned semantics (saf_utility_matrixConv.c:71-96, 174-207; multi
 * :277-298, 368-386): one FFT of numOvrlpAddBlocks * hopSize points (generally not a power of two) per block and channel
 * and a fftSize-long shifting overlap-add buffer, on the general-size device FFT.  Off by default -- the partitioned
 * engine then serves both flags (same causal linear convolution, equal to rounding).  Process-wide switch read by the next
 * saf_matrixConv_create / saf_multiConv_create; also SAFCONV_TRUE_MODE0=1 in the environment.  Handles built this way
 * work with apply / destroy, safconv_last_error[_string], safconv_get_info, safconv_reset_state.
 */
int safconv_set_true_mode0(int enable);

/** Error codes stored per handle / per thread. 0 means OK. */
enum {
    SAFCONV_OK            = 0,
    SAFCONV_ERR_ARG       = 1,  /* invalid argument (sizes <= 0, NULL pointers, TVConv hop > 8192 ...) */
    SAFCONV_ERR_NO_DEVICE = 2,  /* no usable CUDA device / driver */
    SAFCONV_ERR_CUDA      = 3,  /* a CUDA runtime call or kernel failed */
    SAFCONV_ERR_NOMEM     = 4   /* host or device allocation failed */
};

/** Last error of a handle (or of the calling thread's most recent create if h == NULL). */
int         safconv_last_error(void* h);
const char* safconv_last_error_string(void* h);

/** Library build string, e.g. "safconv-b200 0.1 (sm_100a)". */
const char* safconv_version(void);

/** Select the CUDA device used by subsequent `*_create` calls of this thread (default: current device). */
int safconv_set_device(int device);

/**
 * Sharded create: this handle owns output channels [outBegin, outBegin+outCount) of an
 * nCHout-wide problem (output channels are independent: saf_utility_matrixConv.c:218-234).
 * H is still the FULL FLAT nCHout x nCHin x length_h array; apply() then produces
 * outCount x hopSize.  Used for output-channel sharding across GPUs.
 */
void safconv_matrixConv_create_shard(void** const phMC, int hopSize, const float* H, int length_h,
                                     int nCHin, int nCHout, int outBegin, int outCount);

/**
 * As safconv_matrixConv_create_shard, but Hshard holds ONLY this shard's filters
 * (FLAT outCount x nCHin x length_h), so that each process of a one-process-per-GPU job needs
 * just its own part of the filter matrix in host memory.
 */
void safconv_matrixConv_create_from_shard(void** const phMC, int hopSize, const float* Hshard, int length_h,
                                          int nCHin, int nCHout, int outBegin, int outCount);

/** As saf_multiConv_create for channels [chBegin, chBegin+chCount) of H (FLAT nCH x length_h). */
void safconv_multiConv_create_shard(void** const phMC, int hopSize, const float* H, int length_h,
                                    int nCH, int chBegin, int chCount);

/**
 * ONE handle over several GPUs of a box (single host process; csrc/safconv_multi.c).  Output channels are independent
 * in the reference (saf_utility_matrixConv.c:218-234; multiConv channels: .c:388-413), so device g of nDevices owns a
 * contiguous range of output channels (its rows of the filter spectra, its overlap tails and a replica of the
 * frequency-domain delay line); no partial sums cross GPUs.  The handle is used with the UNCHANGED drop-in calls
 * saf_matrixConv_apply / saf_matrixConv_destroy (saf_multiConv_apply / _destroy): host pointers, one block per
 * call, synchronous -- the reference's contract (saf_utility_matrixConv.c:209-235).  A persistent worker thread per
 * device drives that device's shard (look-ahead apply included).  Exchange per block, option "transport":
 *   0 (default)  every device reads the page-locked input block and writes its rows of the page-locked output block
 *                directly over its own PCIe link (pageable caller buffers are staged once): no inter-GPU traffic
 *   1            device 0 uploads the block, ncclBroadcast over NVLink, ncclSend/ncclRecv gather of the output shards
 *                into a channel-major buffer on device 0, one download -- on a per-device side stream.  NCCL
 *                (libnccl.so.2) is loaded with dlopen when this transport is first selected.
 * Further options of a multi-GPU handle: "worker_spin_us" (how long an idle worker spins before it sleeps on a
 * condition variable; default 200), "detect_pinned"; any other option is forwarded to every shard.
 * Environment: SAFCONV_DEVICES="0,1,2,3" | "all" makes the plain saf_matrixConv_create / saf_multiConv_create build
 * such a handle (no source change in the SAF host at all); SAFCONV_MULTI_TRANSPORT, SAFCONV_MULTI_SPIN_US set the defaults.
 * On failure *phMC = NULL and safconv_last_error_string(NULL) says why.  Works with safconv_last_error[_string],
 * safconv_get_info (totals over all devices), safconv_set_option, safconv_reset_state, safconv_synchronize.
 */
void safconv_matrixConv_create_multi(void** const phMC, int hopSize, const float* H, int length_h, int nCHin, int nCHout,
                                     const int* devices, int nDevices);
/** As above for saf_multiConv (H FLAT nCH x length_h): channels are sharded, nothing is replicated. */
void safconv_multiConv_create_multi(void** const phMC, int hopSize, const float* H, int length_h, int nCH,
                                    const int* devices, int nDevices);
/** Devices of a multi-GPU handle: fills devices[0..min(cap, n)) and returns n (0 if h is not such a handle). */
int   safconv_multi_get_devices(void* h, int* devices, int cap);
/** The single-device handle that serves device i of a multi-GPU handle (introspection: safconv_get_info etc.). */
void* safconv_multi_get_shard(void* h, int i);

/**
 * Device-pointer apply: d_in (nCHin x hop) and d_out (nOutLocal x hop) are device
 * pointers on the handle's device.  Enqueues on the handle's stream and returns
 * without synchronising.  Works for matrixConv and multiConv handles.
 */
int safconv_apply_device(void* h, const float* d_in, float* d_out);

/**
 * Enqueue `nBlocks` consecutive blocks: d_in is [nBlocks][nCHin][hop], d_out is
 * [nBlocks][nOutLocal][hop] (device pointers).  No host synchronisation.
 */
int safconv_apply_device_blocks(void* h, const float* d_in, float* d_out, int nBlocks);

/* ========================================================================== */
/*   Drop-in block 2 (reference: saf_utility_fft.h:240-276, real FFT wrapper)  */
/* ========================================================================== */

/**
 * Real <-> half-complex FFT of ANY even size N >= 2, replaces saf_rfft_create / _destroy / _forward / _backward
 * (/root/reference/framework/modules/saf_utilities/saf_utility_fft.h:240-276, .c:531-753; default backend KissFFT:
 * resources/kissFFT/kiss_fftr.c:69-161, kiss_fft.c:93-331).  inputTD / outputTD: N floats; outputFD / inputFD: N/2+1
 * interleaved (re, im) pairs = the reference's float_complex* (declared void* here so that C and C++ callers need no
 * complex type).  Forward unscaled, backward scaled by 1/N and ignoring the imaginary parts of DC and Nyquist.
 * The handle is a resident device plan (radices 4, 2, 3, 5 + any remaining primes; twiddle tables, work arrays,
 * page-locked staging, its own stream): a transform call allocates and uploads nothing.  Host pointers, synchronous.
 * Without a CUDA device *phFFT = NULL (safconv_last_error_string(NULL) says why); forward / backward on NULL are no-ops.
 */
void saf_rfft_create(void** const phFFT, int N);
void saf_rfft_destroy(void** const phFFT);
void saf_rfft_forward(void* const hFFT, float* inputTD, void* outputFD);
void saf_rfft_backward(void* const hFFT, void* inputFD, float* outputTD);

/** Extension: nBatch transforms in one call on a saf_rfft handle (dir 0: in [nBatch][N] real -> out [nBatch][N/2+1] complex;
 *  dir 1: the inverse); the handle's buffers grow to the largest batch seen.  Returns SAFCONV_OK or an error code. */
int safconv_rfft_batch(void* hFFT, int dir, int nBatch, const float* in, float* out);
int         safconv_rfft_last_error(void* hFFT);
const char* safconv_rfft_last_error_string(void* hFFT);
/** Radices of the handle's plan in pass order (fills fac[0..min(cap, n))), returns n. */
int safconv_rfft_get_factors(void* hFFT, int* fac, int cap);

/**
 * Stateless convenience forms (create a plan, transform nBatch signals, destroy): x [nBatch][N] real <-> X [nBatch][N/2+1]
 * interleaved complex, saf_rfft conventions, host pointers, any even N.
 */
int  safconv_rfft_forward(int N, int nBatch, const float* x, float* X);
int  safconv_rfft_backward(int N, int nBatch, const float* X, float* x);

/**
 * Whole-signal per-channel linear convolution, the reference's fftconv / fftfilt
 * (/root/reference/framework/modules/saf_utilities/saf_utility_fft.h:86-91, 107-112; .c:157-228):
 * x FLAT nCH x x_len, h FLAT nCH x h_len (host pointers), y FLAT nCH x (x_len+h_len-1) for fftconv,
 * nCH x x_len (the first x_len samples) for fftfilt.  Served by the multiConv engine (hop-sized blocks through the
 * batched device path) instead of the reference's single nextpow2(x_len+h_len-1)-point FFT: same linear
 * convolution, equal to fp32 rounding.  The safconv_ names return SAFCONV_OK or an error code; `fftconv` and
 * `fftfilt` themselves are exported as WEAK symbols with the reference's void signatures, so the library can be
 * preloaded in front of a SAF build without clashing with a statically linked saf_utility_fft.c.
 */
int  safconv_fftconv(const float* x, const float* h, int x_len, int h_len, int nCH, float* y);
int  safconv_fftfilt(const float* x, const float* h, int x_len, int h_len, int nCH, float* y);
void fftconv(float* x, float* h, int x_len, int h_len, int nCH, float* y);
void fftfilt(float* x, float* h, int x_len, int h_len, int nCH, float* y);

/* ========================================================================== */
/*  Filter producers in front of the convolvers          */
/* ========================================================================== */

/**
 * Convolver from a filter bank that already lies in DEVICE memory: d_H FLAT nCHout x nCHin x length_h on the device the
 * handle is created on (safconv_set_device / current device).  Same handle as saf_matrixConv_create (partitioned engine,
 * hopSize <= 8192); d_H is only read during the call.  This is how the two producers below hand their result over
 * without a device -> host -> device round trip.
 */
void safconv_matrixConv_create_device(void** const phMC, int hopSize, const float* d_H, int length_h, int nCHin, int nCHout);

/**
 * Binaural Ambisonic decoder design on the GPU -- the reference's getBinauralAmbiDecoderMtx / getBinauralAmbiDecoderFilters
 * (/root/reference/framework/modules/saf_hoa/saf_hoa.h:401-471, saf_hoa.c:393-497), same arguments and layouts:
 *   hrtfs          FLAT N_bands x 2 x N_dirs, interleaved (re, im) fp32 (the reference's float_complex*)
 *   hrtf_dirs_deg  FLAT N_dirs x 2, [azimuth, elevation] in degrees
 *   method         BINAURAL_AMBI_DECODER_METHODS (saf_hoa.h:131-171): 0 default (= LS), 1 LS, 2 LSDIFFEQ, 3 SPR, 4 TA, 5 MAGLS
 *   weights        N_dirs integration weights or NULL (uniform 1 / N_dirs)
 *   decMtx         FLAT N_bands x 2 x (order+1)^2 complex;  decFilters FLAT 2 x (order+1)^2 x fftSize real -- the
 *                  nCHout x nCHin x length_h layout saf_matrixConv_create takes (N_bands = fftSize/2 + 1 for the filters)
 * All five designs are built (LS, LSDIFFEQ, SPR, TA, MAGLS), with max-rE weighting and diffuse-field covariance matching.
 * SPR projects on SAF's minimum t-design of degree 2 * order (saf_hoa_internal.c:383-389); those tables are SAF's data and
 * are not copied into this library: it uses what safconv_register_tdesign was given, else SAF's own
 * __HANDLES_Tdesign_dirs_deg / __Tdesign_nPoints_per_degree if the host process carries them (weak references), else the
 * call returns SAFCONV_ERR_ARG and leaves the output untouched -- as does order > 10.
 * itd_s is accepted and, as in the reference (whose TA phase term is exp(0 * itd), saf_hoa_internal.c:494-497), has no
 * influence on the result.  The safconv_ names return SAFCONV_OK or an error code (text: safconv_last_error_string(NULL));
 * the reference's own names are exported as WEAK void functions.
 */
/** Hand a spherical t-design to the SPR decoder: dirs_deg FLAT nPoints x 2, [azimuth, elevation] in degrees (copied;
 *  process-wide, thread-safe).  A SAF host passes __HANDLES_Tdesign_dirs_deg[degree-1], __Tdesign_nPoints_per_degree[degree-1]. */
int  safconv_register_tdesign(int degree, const float* dirs_deg, int nPoints);
int  safconv_getBinauralAmbiDecoderMtx(const void* hrtfs, const float* hrtf_dirs_deg, int N_dirs, int N_bands, int method,
                                       int order, const float* freqVector, const float* itd_s, const float* weights,
                                       int enableDiffCovMatching, int enableMaxReWeighting, void* decMtx);
int  safconv_getBinauralAmbiDecoderFilters(const void* hrtfs, const float* hrtf_dirs_deg, int N_dirs, int fftSize, float fs,
                                           int method, int order, const float* itd_s, const float* weights,
                                           int enableDiffCovMatching, int enableMaxReWeighting, float* decFilters);
void getBinauralAmbiDecoderMtx(void* hrtfs, float* hrtf_dirs_deg, int N_dirs, int N_bands, int method, int order,
                               float* freqVector, float* itd_s, float* weights, int enableDiffCovMatching,
                               int enableMaxReWeighting, void* decMtx);
void getBinauralAmbiDecoderFilters(void* hrtfs, float* hrtf_dirs_deg, int N_dirs, int fftSize, float fs, int method,
                                   int order, float* itd_s, float* weights, int enableDiffCovMatching,
                                   int enableMaxReWeighting, float* decFilters);
/** The same design straight into a (order+1)^2-in, 2-out matrix convolver; the filters never leave the device. */
int  safconv_binauralDecoder_create_matrixConv(void** const phMC, int hopSize, const void* hrtfs, const float* hrtf_dirs_deg,
                                               int N_dirs, int fftSize, float fs, int method, int order, const float* weights,
                                               int enableDiffCovMatching, int enableMaxReWeighting);

/**
 * Shoebox image-source simulator, RIR path -- the reference's ims_shoebox_* API
 * (/root/reference/framework/modules/saf_reverb/saf_reverb.h:93-230, saf_reverb.c:36-295, 541-856): same arguments, same
 * ID assignment, same refresh rules.  ims_shoebox_renderRIRs enumerates the image sources of every pending
 * source / receiver pair on the GPU and accumulates them straight into the RIR taps (tap positions bit-identical to the
 * reference; fractionalDelaysFLAG must be 0, as in the reference).  As in the reference, the band RIRs are summed
 * WITHOUT the octave filterbank (saf_reverb_internal.c:697-702 filters into a scratch buffer that is never read).
 * Receivers up to SH order 10; ims_shoebox_applyEchogramTD (the time-domain renderer) is not part of this path.
 * The reference's names are exported as WEAK symbols that forward to the safconv_ims_shoebox_ ones.
 */
void safconv_ims_shoebox_create(void** phIms, float roomDimensions[3], float* abs_wall, float lowestOctaveBand, int nOctBands,
                                float c_ms, float fs);
void safconv_ims_shoebox_destroy(void** phIms);
void safconv_ims_shoebox_computeEchograms(void* hIms, int maxN, float maxTime_s);
void safconv_ims_shoebox_renderRIRs(void* hIms, int fractionalDelaysFLAG);
void safconv_ims_shoebox_setRoomDimensions(void* hIms, float new_roomDimensions[3]);
void safconv_ims_shoebox_setWallAbsCoeffs(void* hIms, float* abs_wall);
int  safconv_ims_shoebox_addSource(void* hIms, float position_xyz[3], float** pSrc_sig);
int  safconv_ims_shoebox_addReceiverSH(void* hIms, int sh_order, float position_xyz[3], float*** pSH_sigs);
void safconv_ims_shoebox_updateSource(void* hIms, int sourceID, float position_xyz[3]);
void safconv_ims_shoebox_updateReceiver(void* hIms, int receiverID, float position_xyz[3]);
void safconv_ims_shoebox_removeSource(void* hIms, int sourceID);
void safconv_ims_shoebox_removeReceiver(void* hIms, int receiverID);
void ims_shoebox_create(void** phIms, float roomDimensions[3], float* abs_wall, float lowestOctaveBand, int nOctBands, float c_ms, float fs);
void ims_shoebox_destroy(void** phIms);
void ims_shoebox_computeEchograms(void* hIms, int maxN, float maxTime_s);
void ims_shoebox_renderRIRs(void* hIms, int fractionalDelaysFLAG);
void ims_shoebox_setRoomDimensions(void* hIms, float new_roomDimensions[3]);
void ims_shoebox_setWallAbsCoeffs(void* hIms, float* abs_wall);
int  ims_shoebox_addSource(void* hIms, float position_xyz[3], float** pSrc_sig);
int  ims_shoebox_addReceiverSH(void* hIms, int sh_order, float position_xyz[3], float*** pSH_sigs);
void ims_shoebox_updateSource(void* hIms, int sourceID, float position_xyz[3]);
void ims_shoebox_updateReceiver(void* hIms, int receiverID, float position_xyz[3]);
void ims_shoebox_removeSource(void* hIms, int sourceID);
void ims_shoebox_removeReceiver(void* hIms, int receiverID);
/**
 * Accessors (the reference keeps the rendered RIRs inside its handle, ims_scene_data::rirs, without a getter):
 * the RIR of (receiverID, sourceID), FLAT nChannels x length -- a host copy owned by the handle (valid until the pair
 * is rendered again or removed) or the device array itself -- and the number of image sources that went into it.
 */
int  safconv_ims_get_rir(void* hIms, int receiverID, int sourceID, const float** data, int* length, int* nChannels);
int  safconv_ims_get_rir_device(void* hIms, int receiverID, int sourceID, const float** d_data, int* length, int* nChannels);
int  safconv_ims_get_num_images(void* hIms, int receiverID, int sourceID);
/**
 * The rendered RIRs of one receiver as a matrix convolver (configs[3] is such a bank): input channel = source (the
 * active sources in slot order), output channel = SH channel of the receiver, length_h = the longest RIR.  The bank is
 * assembled and transformed on the device.
 */
int  safconv_ims_create_matrixConv(void* hIms, int receiverID, int hopSize, void** const phMC);

/**
 * Offline rendering of a whole signal (BASELINE.json configs[4]): all nFrames blocks are available at once,
 * so the per-bin sum over partitions x inputs becomes a dense contraction that re-uses every filter value
 * for all frames and runs on the tensor cores (tcgen05, fp16 operands split hi+lo with exact power-of-two
 * scaling -- or tf32 with SAFCONV_OFF_KIND=tf32 --, fp32 accumulation).
 * Equal (to fp32 rounding) to a fresh handle's saf_matrixConv_apply called nFrames times:
 * the state before the first frame is zero and the handle's streaming state is neither used nor changed.
 * Layouts are channel-major whole signals: in [nCHin][nFrames*hopSize], out [nOutLocal][nFrames*hopSize].
 * matrixConv handles only; hopSize <= 4096 (more than 64 local outputs are rendered as tiles of 64).
 * _device: device pointers, enqueued on the handle's stream without synchronising.
 */
int safconv_render_offline(void* h, const float* in, float* out, int nFrames);
/* host buffers with nHaloFrames leading history frames: in [nCHin][(nHaloFrames+nFrames)*hopSize], out [nOutLocal][nFrames*hopSize].
 * Both host-buffer calls render signals longer than 512 frames in time segments through a three-stream pipeline
 * (H2D of the next segment and D2H of the previous one beside the kernels; page-locked buffers overlap fully). */
int safconv_render_offline_segment(void* h, const float* in, float* out, int nFrames, int nHaloFrames);
int safconv_render_offline_device(void* h, const float* d_in, float* d_out, int nFrames);
/**
 * One TIME SEGMENT of an offline render (multi-GPU: each GPU renders its own stretch of the signal, no
 * exchange between GPUs).  d_in holds nHaloFrames + nFrames frames per channel: the halo is the audio that
 * precedes the segment (numFilterBlocks frames are enough: the block convolution of frame t reads frames
 * t-P+1..t and the overlap tail of frame t-1; silence for the start of the signal).  Only the nFrames frames
 * of the segment are written: d_out [nOutLocal][nFrames*hopSize].
 */
int safconv_render_offline_segment_device(void* h, const float* d_in, float* d_out, int nFrames, int nHaloFrames);
/** Milliseconds of the last offline render: ms[0] forward FFTs, ms[1] tensor-core GEMM, ms[2] inverse FFTs + overlap-add. */
int safconv_get_offline_times(void* h, float ms[3]);

/** Use an external CUDA stream (cudaStream_t passed as void*) for this handle; NULL restores its own stream. */
int safconv_set_stream(void* h, void* cudaStream);
/** The stream (cudaStream_t as void*) the handle currently enqueues on. */
void* safconv_get_stream(void* h);
/** Block the host until everything enqueued by this handle has finished. */
int safconv_synchronize(void* h);

/** Reset the convolver state (delay line + overlap tails) to zero without re-transforming the filters. */
int safconv_reset_state(void* h);

/** Introspection of a handle's plan (any out-pointer may be NULL). */
typedef struct safconv_info {
    int kind;              /* 0 matrix, 1 multi, 2 time-varying */
    int hopSize, length_h, nCHin, nCHout, nOutLocal, outBegin;
    int fftSize, nBinsPacked, numFilterBlocks;
    int macGrid, macStages, macThreads;     /* launch geometry of the filter-streaming MAC */
    int maxBatch;                           /* blocks per launch group of safconv_apply_device_blocks */
    int device;
    size_t bytesFilters, bytesDelayLine;    /* resident device bytes */
    double algBytesPerBlock;                /* SURVEY.md §8(d) algorithmic bytes per block for this handle */
    double macAlgBytesPerBlock;             /* H + delay-line bytes read by the MAC per block */
} safconv_info;
int safconv_get_info(void* h, safconv_info* info);

/**
 * Timing helper for benchmarks.  safconv_enable_kernel_timing(h, nGroups) allocates CUDA events for up to
 * nGroups launch groups (0 disables).  A launch group is what one saf_*_apply call or one batch of
 * safconv_apply_device_blocks enqueues: [forward FFT launch][ONE filter-streaming MAC launch covering all
 * blocks of the group][inverse FFT + overlap-add launch(es)]; while enabled, events are recorded between
 * those launches on the handle's stream.
 * safconv_get_kernel_totals synchronises the stream and returns the TOTAL milliseconds per kernel class
 * (msTotal[0] forward FFT, [1] MAC [or the fused multiConv kernel], [2] inverse FFT + overlap-add), the number
 * of launch groups and the number of blocks they covered, and restarts the recording.
 * safconv_get_kernel_times is the same divided by the number of blocks (average per block).
 */
int safconv_enable_kernel_timing(void* h, int nGroups);
int safconv_get_kernel_totals(void* h, float msTotal[3], int* nLaunchGroups, int* nBlocks);
int safconv_get_kernel_times(void* h, float ms[3], int* nBlocksAveraged);

/** Tuning knobs (mostly for benchmarks / tests). Returns 0 on success.
 *    "mac_hints"    L2 eviction-priority hints of the MAC: 0 none, 1 filters evict_first (streamed once per block),
 *                   2 filters evict_last (a filter set that fits in L2); default: 2 up to 64 MB of spectra, else 1
 *    "lookahead"    0/1 (default 1 for matrix convolvers with more than one partition): saf_matrixConv_apply returns
 *                   when the block's output is on the host while the GPU already accumulates the partitions of the
 *                   NEXT block that do not depend on it; 0 = plain forward FFT -> MAC -> inverse FFT per call
 *    "use_graph"    0/1 replay the per-block launch sequence of saf_*_apply from a CUDA graph
 *    "batching"     0/1 (default 1) safconv_apply_device_blocks shares one forward-FFT and one inverse-FFT
 *                   launch across the blocks handed over together (outputs are bit-identical either way)
 *    "small_fused"  0/1 (default 1) saf_matrixConv_apply of a small problem (few inputs x outputs, filter spectra
 *                   that stay in L2) runs as ONE fused kernel that reads / writes the page-locked host buffers
 *                   directly: one launch + one synchronisation per block (real-time latency path); also gates the other
 *                   zero-copy host paths (multiConv single launch, single-partition matrix convolvers)
 *    "detect_pinned" 0/1 (default 1) saf_*_apply copies straight from/to caller buffers that are already
 *                   page-locked instead of going through the handle's own pinned staging buffers
 *    "flag_wait"    0/1 (default 1) the one-launch latency kernel writes the call's sequence number into a page-locked word
 *                   behind its last output store and saf_matrixConv_apply polls that word instead of synchronising the
 *                   stream (falls back to the stream after 2 ms, which also reports errors)
 *    "resident_us"  N > 0: RESIDENT latency kernel for small matrix problems (the ones "small_fused" serves with the
 *                   thread-block-cluster kernel: FFT size 128 .. 2048).  The kernel stays on 8 SMs and serves one block per
 *                   doorbell -- saf_matrixConv_apply writes the buffer addresses + a sequence number into a page-locked
 *                   mailbox and polls the completion word: no kernel launch per call (configs[1]: p50 20 -> 15 us).  It
 *                   leaves by itself after N microseconds without a call (the next call starts it again), and every other
 *                   call on the handle (device-pointer apply, reset, options, destroy) stops it first.  While it polls,
 *                   the handle's stream is busy: cudaDeviceSynchronize() in the host blocks for up to N us.  Default 0 (off);
 *                   SAFCONV_RESIDENT_US sets the default.  Results are bit-identical to the one-launch kernel.
 * Environment variables read at create / first use (benchmark and debugging knobs, all optional):
 *    SAFCONV_LOOKAHEAD, SAFCONV_SMALL_FUSED, SAFCONV_MAC_HINTS, SAFCONV_MAC_STAGES, SAFCONV_MAC_STAGE_KB, SAFCONV_MAX_BATCH,
 *    SAFCONV_MULTI_WFFT, SAFCONV_TRACE / SAFCONV_TIMELINE=n (device timeline of the look-ahead apply on stderr),
 *    SAFCONV_LA_DEPTH (1 | 2: tail passes queued ahead), SAFCONV_HEAD_IN_K3, SAFCONV_TAIL_RESERVE_SMS (default 32: SMs a
 *    tail pass leaves to the kernels of the caller's critical path), SAFCONV_SMALL_CLUSTER, SAFCONV_FLAG_WAIT,
 *    SAFCONV_HOSTTRACE, SAFCONV_KSTAMPS (host / in-kernel phase times of the latency kernel);
 *    offline path: SAFCONV_OFF_KIND (f16 | tf32), SAFCONV_OFF_NT, SAFCONV_OFF_WFFT, SAFCONV_OFF_FLUSH, SAFCONV_OFF_PIPELINE,
 *    SAFCONV_OFF_FPC, SAFCONV_OFF_OPC, SAFCONV_OFF_THREADS.
 */
int safconv_set_option(void* h, const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* SAFCONV_B200_H_INCLUDED */
