// ntt32_host.inl -- host planner + C ABI for the 32-bit field (included by gsn_lib.cu)
extern "C" {
int gsn_ntt32_host(gsn_ctx *, uint32_t *, size_t, uint32_t, uint32_t, int) { return fail(GSN_ERR_INVALID_ARG, "ntt32: not built yet"); }
int gsn_ntt32_device(gsn_ctx *, uint32_t *, size_t, size_t, uint32_t, uint32_t, int, void *) { return fail(GSN_ERR_INVALID_ARG, "ntt32: not built yet"); }
int gsn_ntt32_time_device(gsn_ctx *, uint32_t *, size_t, size_t, uint32_t, uint32_t, int, int, float *) { return fail(GSN_ERR_INVALID_ARG, "ntt32: not built yet"); }
}
