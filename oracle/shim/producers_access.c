/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 *
 * The reference's image-source simulator keeps its rendered room impulse responses inside the scene handle
 * (ims_scene_data::rirs, /root/reference/framework/modules/saf_reverb/saf_reverb_internal.h:200-260) and offers no
 * public accessor (saf_reverb.h:93-230).  This translation unit is compiled TOGETHER with the unmodified reference
 * sources into oracle/_ref/libsaf_ref_producers.so and only reads that struct, so that tests / golden generation can
 * compare the rendered RIRs and the echogram that produced them.
 */
#include "saf_reverb.h"
#include "saf_reverb_internal.h"
#include "saf_utility_loudspeaker_presets.h"

static int find_src(ims_scene_data* sc, int id)
{ for (int i = 0; i < IMS_MAX_NUM_SOURCES; i++) if (sc->srcs[i].ID == id) return i; return -1; }
static int find_rec(ims_scene_data* sc, int id)
{ for (int i = 0; i < IMS_MAX_NUM_RECEIVERS; i++) if (sc->recs[i].ID == id) return i; return -1; }

/* rendered RIR of (receiverID, sourceID): data FLAT nChannels x length; returns 0 on success */
int oracle_ims_get_rir(void* hIms, int receiverID, int sourceID, float** data, int* length, int* nChannels)
{
    ims_scene_data* sc = (ims_scene_data*)hIms;
    const int r = find_rec(sc, receiverID), s = find_src(sc, sourceID);
    if (r < 0 || s < 0) return -1;
    *data = sc->rirs[r][s].data; *length = sc->rirs[r][s].length; *nChannels = sc->rirs[r][s].nChannels;
    return 0;
}

/* number of image sources of the pair's echogram; optionally copies the sorted propagation times (seconds) */
int oracle_ims_get_echogram(void* hIms, int receiverID, int sourceID, float* times, int cap)
{
    ims_scene_data* sc = (ims_scene_data*)hIms;
    const int r = find_rec(sc, receiverID), s = find_src(sc, sourceID);
    if (r < 0 || s < 0) return -1;
    ims_core_workspace* w = (ims_core_workspace*)sc->hCoreWrkSpc[r][s];
    echogram_data* e = (echogram_data*)w->hEchogram_abs[0];
    if (times) for (int i = 0; i < e->numImageSources && i < cap; i++) times[i] = e->time[i];
    return e->numImageSources;
}

/* the reference's minimum t-design of a given degree (1..21), [azimuth, elevation] in degrees
 * (saf_utility_loudspeaker_presets.h:286-301): the SPR decoder projects on the one of degree 2 * order
 * (saf_hoa_internal.c:383-389).  Tests hand it to the product through safconv_register_tdesign. */
const float* oracle_tdesign(int degree, int* nPoints)
{
    if (degree < 1 || degree > 21) return 0;
    *nPoints = __Tdesign_nPoints_per_degree[degree - 1];
    return __HANDLES_Tdesign_dirs_deg[degree - 1];
}
