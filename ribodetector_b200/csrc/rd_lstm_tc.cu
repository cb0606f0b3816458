// rd_lstm_tc.cu — K2 (RD_PREC_TC_EXACT / RD_PREC_TC_FAST): tcgen05 forward LSTM.  (placeholder
// until the tensor-core kernel lands; fails loudly, never falls back)
#include "rd_common.cuh"

int rd_tc_create(rd_handle*, const float*, const float*) { return RD_OK; }
void rd_tc_destroy(rd_handle*) {}
int rd_launch_lstm_tc(rd_handle* h, int64_t, int, int, float*, cudaStream_t) {
    h->err = "tensor-core precision modes are not built in this library";
    return RD_ERR_UNSUPPORTED;
}
