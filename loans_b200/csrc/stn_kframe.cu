// stn_kframe.cu -- backward of axis-aligned crops with SEVERAL crops per frame (BASELINE config 4: 16 jittered boxes per
// frame; no Chainer equivalent, semantics in include/loans_stn.h).  Placeholder until the kernel lands: declines every call.
#include "stn_common.cuh"

namespace stn {

int launch_crop_bwd_kframe(CropParams p, int gy_dtype, cudaStream_t stream)
{
    (void)p; (void)gy_dtype; (void)stream;
    return -1;
}

}  // namespace stn
