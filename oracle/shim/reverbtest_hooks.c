/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 *
 * The reference's own unit test of the image-source simulator, test__ims_shoebox_RIR
 * (/root/reference/test/src/test__reverb_module.c:27-96), is compiled UNMODIFIED into
 * oracle/_ref/libsaf_ref_reverbtest_b200.so with its ims_shoebox_* calls resolved from libsafconv_b200.so.  The test
 * asserts nothing and destroys its scene at the end, so the translation unit is compiled with
 * -Dims_shoebox_destroy=reverbtest_capture_destroy: the hook below copies the rendered RIRs of every live
 * source / receiver pair out of the product's handle (safconv_ims_get_rir) before it forwards to the real destroy.
 * tests/test_example_cores.py compares them with the golden outputs of the compiled reference for the same sequence.
 *
 * ims_shoebox_applyEchogramTD (used by the second test function of that file, which is never called) is not part of the
 * product's RIR path; a stub satisfies the dynamic linker.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern int  safconv_ims_get_rir(void* hIms, int receiverID, int sourceID, const float** data, int* length, int* nChannels);
extern void safconv_ims_shoebox_destroy(void** phIms);

#define RT_MAX 16
static struct { int rid, sid, len, nch; float* data; } g_cap[RT_MAX * RT_MAX];
static int g_ncap = 0;

void reverbtest_capture_destroy(void** phIms)
{
    for (int i = 0; i < g_ncap; i++) free(g_cap[i].data);
    g_ncap = 0;
    for (int rid = 0; rid < 4; rid++)
        for (int sid = 0; sid < RT_MAX; sid++) {
            const float* d = NULL; int len = 0, nch = 0;
            if (safconv_ims_get_rir(*phIms, rid, sid, &d, &len, &nch) != 0 || !d) continue;
            g_cap[g_ncap].rid = rid; g_cap[g_ncap].sid = sid; g_cap[g_ncap].len = len; g_cap[g_ncap].nch = nch;
            g_cap[g_ncap].data = (float*)malloc(sizeof(float) * (size_t)len * nch);
            memcpy(g_cap[g_ncap].data, d, sizeof(float) * (size_t)len * nch);
            g_ncap++;
        }
    safconv_ims_shoebox_destroy(phIms);
}

int reverbtest_num_captured(void) { return g_ncap; }
int reverbtest_get(int i, int* rid, int* sid, int* len, int* nch, const float** data)
{
    if (i < 0 || i >= g_ncap) return -1;
    *rid = g_cap[i].rid; *sid = g_cap[i].sid; *len = g_cap[i].len; *nch = g_cap[i].nch; *data = g_cap[i].data;
    return 0;
}

void ims_shoebox_applyEchogramTD(void* hIms, long receiverID, int nSamples, int fractionalDelaysFLAG)
{
    (void)hIms; (void)receiverID; (void)nSamples; (void)fractionalDelaysFLAG;
    fprintf(stderr, "ims_shoebox_applyEchogramTD is not part of libsafconv_b200's RIR path\n");
    abort();
}
