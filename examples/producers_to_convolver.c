/*
 * examples/producers_to_convolver.c -- the two filter producers feeding a convolver, through the C ABI only
 * (include/safconv_b200.h).  Build:  gcc -std=c99 -Iinclude examples/producers_to_convolver.c \
 *                                        -Lspatial_audio_framework_b200 -lsafconv_b200 -Wl,-rpath,$PWD/spatial_audio_framework_b200 -lm
 *
 * (1) a shoebox room with 4 sources and one 3rd-order Ambisonic receiver -> 16 x 4 room impulse responses -> matrix
 *     convolver, bank assembled and transformed on the device (safconv_ims_create_matrixConv);
 * (2) a binaural decoder for those 16 Ambisonic channels from a (here: synthetic) HRTF set -> 2 x 16 FIRs -> second
 *     convolver (safconv_binauralDecoder_create_matrixConv);
 * (3) one block of 4 source signals through both: sources -> Ambisonics -> ears.
 * The calls are the reference's own (ims_shoebox_*, saf_matrixConv_apply: saf_reverb.h:93-230,
 * saf_utility_matrixConv.h:82-86) plus the two hand-over functions.  Without a CUDA device every create fails with an
 * error string and the program says so (exit code 2): the library has no CPU path.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "safconv_b200.h"

#define HOP 512

int main(void)
{
    /* (1) room -> RIR bank -> convolver */
    float room[3] = {10.0f, 7.0f, 3.0f};
    float abs_wall[2][6] = {{0.20f, 0.22f, 0.15f, 0.25f, 0.21f, 0.24f}, {0.35f, 0.40f, 0.27f, 0.45f, 0.42f, 0.48f}};
    float src[4][3] = {{5.1f, 6.0f, 1.1f}, {2.1f, 1.0f, 1.3f}, {4.4f, 3.0f, 1.4f}, {6.4f, 4.0f, 1.3f}};
    float rec[3] = {8.8f, 5.5f, 0.9f};
    void* hIms = NULL;
    ims_shoebox_create(&hIms, room, &abs_wall[0][0], 125.0f, 2, 343.0f, 48e3f);
    if (!hIms) { fprintf(stderr, "ims_shoebox_create: %s\n", safconv_last_error_string(NULL)); return 2; }
    for (int i = 0; i < 4; i++) ims_shoebox_addSource(hIms, src[i], NULL);
    const int recID = ims_shoebox_addReceiverSH(hIms, 3, rec, NULL);
    ims_shoebox_computeEchograms(hIms, -1, 0.25f);          /* 0.25 s of reflections */
    ims_shoebox_renderRIRs(hIms, 0);
    const float* rir = NULL; int len = 0, nch = 0;
    if (safconv_ims_get_rir(hIms, recID, 0, &rir, &len, &nch)) { fprintf(stderr, "%s\n", safconv_last_error_string(NULL)); return 1; }
    printf("room: %d x %d taps per source, %d image sources for source 0\n", nch, len, safconv_ims_get_num_images(hIms, recID, 0));
    void* hRoom = NULL;
    if (safconv_ims_create_matrixConv(hIms, recID, HOP, &hRoom)) { fprintf(stderr, "%s\n", safconv_last_error_string(NULL)); return 1; }

    /* (2) HRTFs -> decoder filters -> convolver (a real host reads a SOFA file here) */
    enum { ND = 240, FFT = 512, NB = FFT / 2 + 1, ORDER = 3, NSH = 16 };
    float* hrtfs = (float*)malloc(sizeof(float) * 2 * NB * 2 * ND);      /* NB x 2 x ND complex */
    float* dirs = (float*)malloc(sizeof(float) * 2 * ND);
    for (int d = 0; d < ND; d++) {                                        /* Fibonacci grid, head-like delays and levels */
        const double z = 1.0 - 2.0 * (d + 0.5) / ND, az = fmod(3.14159265358979 * (1.0 + sqrt(5.0)) * (d + 0.5), 6.28318530717959) - 3.14159265358979;
        dirs[2 * d] = (float)(az * 180.0 / 3.14159265358979); dirs[2 * d + 1] = (float)(asin(z) * 180.0 / 3.14159265358979);
        const double uy = sqrt(1.0 - z * z) * sin(az);
        for (int e = 0; e < 2; e++) {
            const double s = e ? -1.0 : 1.0, tau = 0.0003 - s * (0.0875 / 343.0) * uy / 2.0, g = 1.0 + 0.4 * s * uy;
            for (int k = 0; k < NB; k++) {
                const double f = k * 48000.0 / FFT, ph = -6.28318530717959 * f * tau, a = g / (1.0 + (f / 16000.0) * (f / 16000.0));
                hrtfs[2 * (((size_t)k * 2 + e) * ND + d)] = (float)(a * cos(ph));
                hrtfs[2 * (((size_t)k * 2 + e) * ND + d) + 1] = (float)(a * sin(ph));
            }
        }
    }
    void* hEars = NULL;
    if (safconv_binauralDecoder_create_matrixConv(&hEars, HOP, hrtfs, dirs, ND, FFT, 48000.0f, 5 /* MAGLS */, ORDER, NULL, 1, 1)) {
        fprintf(stderr, "%s\n", safconv_last_error_string(NULL)); return 1;
    }

    /* (3) one block: 4 sources -> 16 Ambisonic channels -> 2 ears */
    static float in[4 * HOP], ambi[NSH * HOP], ears[2 * HOP];
    in[0] = 1.0f;                                                         /* a click on source 0 */
    saf_matrixConv_apply(hRoom, in, ambi);
    saf_matrixConv_apply(hEars, ambi, ears);
    double e = 0.0;
    for (int i = 0; i < 2 * HOP; i++) e += (double)ears[i] * ears[i];
    printf("first block: energy at the ears %.6g\n", e);

    saf_matrixConv_destroy(&hEars);
    saf_matrixConv_destroy(&hRoom);
    ims_shoebox_destroy(&hIms);
    free(hrtfs); free(dirs);
    return 0;
}
