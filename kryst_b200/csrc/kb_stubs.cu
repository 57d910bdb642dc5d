// kb_stubs.cu — entry points not implemented yet (fail loudly with KB_UNSUPPORTED).
#include "kb_objects.h"
#define KB_NYI(name) do { kb_set_error(name ": not implemented yet"); return KB_UNSUPPORTED; } while (0)
int kb_csr_build_dist(kb_csr_s*) { KB_NYI("kb_csr_create_dist"); }
int kb_halo_exchange(kb_csr_s* A, double*) { if (!A->dist || A->ctx->size == 1) return KB_OK; KB_NYI("halo exchange"); }
int kb_allreduce_slots(kb_ctx_s* c, double*, int) { if (c->size == 1) return KB_OK; KB_NYI("allreduce"); }
void kb_halo_free(KbHalo*) {}
int kb_comm_destroy_internal(kb_ctx_s*) { return KB_OK; }
void kb_gmres_ws_free(KbGmresWs*) {}
int kb_ilu0_build(kb_pc_s*) { KB_NYI("ilu0"); }
void kb_ilu0_free(kb_pc_s*) {}
int kb_ilu0_apply_dev(kb_pc_s*, const double*, double*) { KB_NYI("ilu0 apply"); }
extern "C" {
int kb_comm_unique_id(void*) { KB_NYI("kb_comm_unique_id"); }
int kb_comm_init(kb_ctx, int, int, const void*) { KB_NYI("kb_comm_init"); }
int kb_comm_rank(kb_ctx c) { return c->rank; }
int kb_comm_size(kb_ctx c) { return c->size; }
int kb_comm_barrier(kb_ctx c) { if (c->size == 1) return KB_OK; KB_NYI("kb_comm_barrier"); }
int kb_comm_all_reduce(kb_ctx c, double local, double* global) { if (c->size == 1) { *global = local; return KB_OK; } KB_NYI("kb_comm_all_reduce"); }
int kb_pc_create_ilu0(kb_csr, kb_pc* out) { *out = nullptr; KB_NYI("kb_pc_create_ilu0"); }
int kb_pc_ilu0_get_factors(kb_pc, double*, uint64_t*) { KB_NYI("kb_pc_ilu0_get_factors"); }
int kb_pc_ilu0_get_levels(kb_pc, int, uint64_t*, uint64_t*, uint64_t*) { KB_NYI("kb_pc_ilu0_get_levels"); }
int kb_gmres_solve(kb_csr, kb_pc, const double*, double*, uint64_t, double, uint64_t, int, uint32_t, kb_stats*) { KB_NYI("kb_gmres_solve"); }
}
