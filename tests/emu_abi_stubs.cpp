// TEST-ONLY: see tests/emu_abi.cpp.  C-ABI entry points the CPU emulation does not provide; they exist so that the Python binding,
// which resolves every symbol of include/gubernator_b200.h when it loads a library, can load the emulated one.
extern "C" int emu_abi_not_emulated(const char* name);
#define NOT_EMULATED(name) extern "C" int name() { return emu_abi_not_emulated(#name); }
NOT_EMULATED(gub_submit_device) NOT_EMULATED(gub_submit_device_n) NOT_EMULATED(gub_set_profiling) NOT_EMULATED(gub_get_profile)
NOT_EMULATED(gub_hash_keys_device) NOT_EMULATED(gub_route_device) NOT_EMULATED(gub_unroute_device) NOT_EMULATED(gub_gq_create)
NOT_EMULATED(gub_gq_accumulate_device) NOT_EMULATED(gub_gq_drain_device) NOT_EMULATED(gub_make_updates_device) NOT_EMULATED(gub_add_items_device)
NOT_EMULATED(gub_route_owner_device) NOT_EMULATED(gub_route_global_device) NOT_EMULATED(gub_p2p_create) NOT_EMULATED(gub_p2p_export)
NOT_EMULATED(gub_p2p_connect) NOT_EMULATED(gub_p2p_connect_local) NOT_EMULATED(gub_p2p_step) NOT_EMULATED(gub_p2p_step_streams)
NOT_EMULATED(gub_p2p_status) NOT_EMULATED(gub_p2p_enable_global) NOT_EMULATED(gub_nccl_unique_id) NOT_EMULATED(gub_p2p_nccl_init) NOT_EMULATED(gub_p2p_nccl_init_local)
NOT_EMULATED(gub_global_tick) NOT_EMULATED(gub_global_tick_local_all) NOT_EMULATED(gub_p2p_step_local_all) NOT_EMULATED(gub_gq_dropped) NOT_EMULATED(gub_set_sweep)
NOT_EMULATED(gub_set_trace) NOT_EMULATED(gub_get_trace) NOT_EMULATED(gub_get_trace_raw) NOT_EMULATED(gub_get_ktrace) NOT_EMULATED(gub_keys_layout) NOT_EMULATED(gub_submit_keys_async)
extern "C" void gub_gq_destroy() {}
extern "C" void gub_p2p_destroy() {}
