// Temporary: entry points not implemented yet return COOPCAP_ERR_UNSUPPORTED.
#include "../../include/coopcap.h"
#include "common.cuh"
#define STUB(name) coopcap::set_last_error(#name ": not implemented yet"); return coopcap::CC_ERR_UNSUPPORTED;
extern "C" {
int coopcap_st_backward(const coopcap_speaker*, const void*, const void*, float*, void*, coopcap_stream_t) { STUB(st_backward) }
int coopcap_logp_backward(const coopcap_speaker*, const int64_t*, const float*, void*, coopcap_stream_t) { STUB(logp_backward) }
int coopcap_speaker_decode_bwd(const coopcap_speaker*, const coopcap_speaker_grads*, coopcap_stream_t) { STUB(decode_bwd) }
int coopcap_listener_pack_weights(const coopcap_listener_pack*, coopcap_stream_t) { STUB(listener_pack) }
int coopcap_listener_fwd(const coopcap_listener*, coopcap_stream_t) { STUB(listener_fwd) }
int coopcap_listener_bwd(const coopcap_listener*, const coopcap_listener_grads*, coopcap_stream_t) { STUB(listener_bwd) }
int coopcap_clamp_adam(float*, const float*, float*, float*, int64_t, float, float, float, float, float, float, float, int, coopcap_stream_t) { STUB(clamp_adam) }
}
