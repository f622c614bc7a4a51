// stubs.cu — verbs of include/meepo.h that are not implemented yet in the CUDA build.
// They fail loudly (MEEPO_EINVAL + message); nothing falls back to a CPU path.
#include "table.h"
using namespace meepo;
#define NOT_YET(name) return fail(MEEPO_EINVAL, name ": not implemented in the CUDA build yet")
extern "C" {
MEEPO_API meepo_status meepo_evict(meepo_table*, int32_t, double, uint64_t*, void*) { NOT_YET("meepo_evict"); }
MEEPO_API meepo_status meepo_spill_readmit(meepo_table*, const uint64_t*, uint64_t, uint8_t*) { NOT_YET("meepo_spill_readmit"); }
MEEPO_API meepo_status meepo_export_buffers(meepo_table*, uint64_t*, void*, void*, uint64_t*, uint32_t*, uint64_t, uint64_t*) { NOT_YET("meepo_export_buffers"); }
MEEPO_API meepo_status meepo_import_buffers(meepo_table*, const uint64_t*, const void*, const void*, const uint64_t*, const uint32_t*, uint64_t, uint8_t*) { NOT_YET("meepo_import_buffers"); }
MEEPO_API meepo_status meepo_export(meepo_table*, const char*) { NOT_YET("meepo_export"); }
MEEPO_API meepo_status meepo_import(meepo_table*, const char*) { NOT_YET("meepo_import"); }
}
