// sph_multi.cu -- slab-decomposed multi-GPU step (placeholder until the single-GPU path is proven).
#include "sph_context.h"

namespace sphb200 {
int multi_step(SphContext* c, float) { return fail(c, SPH_ERR_UNSUPPORTED, "slab mode not built yet"); }
void multi_teardown(SphContext*) {}
}

using namespace sphb200;
extern "C" {
size_t sph_comm_id_bytes(void) { return 128; }
int sph_comm_get_id(void*, size_t) { return SPH_ERR_UNSUPPORTED; }
int sph_comm_init(SphContext* c, int, int, const void*, size_t) { return fail(c, SPH_ERR_UNSUPPORTED, "slab mode not built yet"); }
int sph_comm_set_planes(SphContext* c, const float*) { return fail(c, SPH_ERR_UNSUPPORTED, "slab mode not built yet"); }
int sph_upload_owned(SphContext* c, uint32_t, const uint32_t*, const float*, const float*) { return fail(c, SPH_ERR_UNSUPPORTED, "slab mode not built yet"); }
int sph_download_owned(SphContext* c, int, uint32_t*, void*, size_t, uint32_t*) { return fail(c, SPH_ERR_UNSUPPORTED, "slab mode not built yet"); }
int sph_comm_stats(const SphContext*, uint32_t*) { return SPH_ERR_UNSUPPORTED; }
}
