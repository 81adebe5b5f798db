// Entry points shared by the engine (swb_engine.cu) and the engine group (swb_group.cu) that are not part of the C ABI.
#pragma once
#include <stdint.h>
#include "../../include/swb.h"

// swb_db_load with the length sort already done: sorted_order = the ids of the whole database by descending length
// (swb_sort_by_length, swb_plan.h), shared by all the engines of a group; NULL sorts here
int swb_db_load_sorted(swb_engine *e, const uint8_t *codes, const uint64_t *offsets, uint32_t n, uint32_t shard,
                       uint32_t nshards, const uint32_t *sorted_order);
// swb_search_batch_scatter with a row map: the vector of query q goes to row row_of[q] of scores_full (NULL: row q)
int swb_search_batch_rows(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                          int32_t *scores_full, uint64_t n_total, const uint32_t *row_of);
