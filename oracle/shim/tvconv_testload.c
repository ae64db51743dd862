/* TEST INFRASTRUCTURE ONLY.
 * The reference's tvconv plug-in core (examples/src/tvconv/tvconv.c) can only load impulse responses from a SOFA file
 * (tvconv_setFiltersAndPositions, tvconv.c:243-263, needs SAF_ENABLE_SOFA_READER_MODULE = netCDF/zlib, absent here).
 * This shim does what that branch does with the fields of tvconv_data -- IRs, listener positions, channel counts --
 * from caller-supplied arrays, so that tests/test_example_cores.py can drive the UNMODIFIED core
 * (tvconv_init / tvconv_setTargetPosition / tvconv_process -> saf_TVConv_create / saf_TVConv_apply). */
#include "tvconv_internal.h"
#include <string.h>

void tvconv_testload(void* const hTVCnv, const float* irs, int nPositions, int nChannels, int irLength, int fs,
                     const float* positions /* nPositions x 3 */)
{
    tvconv_data* pData = (tvconv_data*)hTVCnv;
    int i;
    pData->ir_fs = fs;
    pData->ir_length = irLength;
    pData->nIrChannels = nChannels;
    pData->nListenerPositions = nPositions;
    pData->irs = (float**)realloc2d((void**)pData->irs, nPositions, nChannels * irLength, sizeof(float));
    for (i = 0; i < nPositions; i++)
        memcpy(pData->irs[i], irs + (size_t)i * nChannels * irLength, (size_t)nChannels * irLength * sizeof(float));
    pData->listenerPositions = (vectorND*)realloc1d((void*)pData->listenerPositions, nPositions * sizeof(vectorND));
    memcpy(pData->listenerPositions, positions, nPositions * sizeof(vectorND));
    pData->nOutputChannels = SAF_MIN(pData->nIrChannels, MAX_NUM_CHANNELS);
    tvconv_setMinMaxDimensions(hTVCnv);
    pData->position_idx = 0;
    pData->codecStatus = CODEC_STATUS_INITIALISED;
    pData->reInitFilters = 1;
}
