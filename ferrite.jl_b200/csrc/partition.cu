// Partitioned multi-GPU assembly (placeholder until the partition plan lands).
#include "common.h"

struct fb2_part { int dummy; };

#define FB2_NYI(name) return fb2_fail(FB2_ERR_UNSUPPORTED, name ": not implemented yet")

extern "C" int fb2_partition_create(fb2_dh*, fb2_pattern*, int, int, fb2_part**) { FB2_NYI("fb2_partition_create"); }
extern "C" int fb2_partition_info(fb2_part*, int64_t*, int64_t*, int64_t*, int64_t*) { FB2_NYI("fb2_partition_info"); }
extern "C" int fb2_partition_cells(fb2_part*, int64_t*) { FB2_NYI("fb2_partition_cells"); }
extern "C" int fb2_partition_peer_counts(fb2_part*, int, int64_t*, int64_t*, int64_t*, int64_t*) { FB2_NYI("fb2_partition_peer_counts"); }
extern "C" int fb2_assembler_set_partition(fb2_assembler*, fb2_part*) { FB2_NYI("fb2_assembler_set_partition"); }
extern "C" int fb2_partition_pack(fb2_part*, int, const double*, const double*, double*) { FB2_NYI("fb2_partition_pack"); }
extern "C" int fb2_partition_unpack_add(fb2_part*, int, const double*, double*, double*) { FB2_NYI("fb2_partition_unpack_add"); }
extern "C" int fb2_partition_mask_unowned(fb2_part*, double*, double*) { FB2_NYI("fb2_partition_mask_unowned"); }
extern "C" int fb2_partition_destroy(fb2_part*) { return FB2_OK; }
extern "C" int fb2_comm_unique_id(void*) { FB2_NYI("fb2_comm_unique_id"); }
extern "C" int fb2_comm_init_rank(fb2_ctx*, const void*, int, int) { FB2_NYI("fb2_comm_init_rank"); }
extern "C" int fb2_comm_destroy(fb2_ctx*) { return FB2_OK; }
extern "C" int fb2_assemble_distributed(fb2_assembler*, fb2_part*, int, const void*, size_t, const double*, double*, double*,
                                        const fb2_asm_opts*) { FB2_NYI("fb2_assemble_distributed"); }
