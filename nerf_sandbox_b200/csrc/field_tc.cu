// K1 (bf16 tensor-core mode) -- placeholder until the tcgen05 chain lands; every entry fails loudly.
#include "nsb_common.cuh"
namespace nsb {
size_t tc_packed_bytes() { return 0; }
size_t tc_workspace_bytes(int64_t, int) { return 256; }
int tc_pack(const float*, void*, cudaStream_t) { return NSB_OK; }
int tc_field_fwd_rays(const float*, const float*, const float*, const float*, const float*, const void*, float*, void*,
                      int64_t, int, int, cudaStream_t) { return NSB_E_BADARG; }
int tc_field_fwd_enc(const float*, const float*, const void*, float*, void*, int64_t, int, cudaStream_t) { return NSB_E_BADARG; }
int tc_field_bwd(const float*, const void*, float*, void*, int64_t, cudaStream_t) { return NSB_E_BADARG; }
}  // namespace nsb
