// ops_pending.cu -- entry points whose kernels are not written yet fail loudly.
#include "core.hpp"
using namespace wdb;
extern "C" {
int wdb_column_minmax(int, void *, const wdb_col_t *, double *, double *) { return fail("wdb_column_minmax: not implemented yet"); }
int wdb_multi_project_filter_host(int, const wdb_col_t *, int, const char *, const char *, float *, int64_t, int, int64_t *) { return fail("wdb_multi_project_filter_host: not implemented yet"); }
int wdb_agg_create(int, int64_t, wdb_agg_t **) { return fail("wdb_agg_create: not implemented yet"); }
int wdb_agg_destroy(wdb_agg_t *) { return fail("not implemented yet"); }
int wdb_agg_reset(wdb_agg_t *, void *) { return fail("not implemented yet"); }
int wdb_agg_consume(wdb_agg_t *, void *, const wdb_col_t *, int, const char *, const char *, const char *, int64_t, int64_t) { return fail("not implemented yet"); }
int wdb_agg_merge(wdb_agg_t *, void *, const int32_t *, const double *, const int64_t *, const double *, const double *, const int64_t *, int64_t) { return fail("not implemented yet"); }
int wdb_agg_size(wdb_agg_t *, void *, int64_t *) { return fail("not implemented yet"); }
int wdb_agg_export(wdb_agg_t *, void *, int, int, int32_t *, float *, double *, int64_t *, double *, double *, int64_t *, int64_t, int64_t *) { return fail("not implemented yet"); }
int wdb_group_agg(int, void *, const wdb_col_t *, int, const char *, const char *, const char *, int, int, int64_t, int64_t, int32_t *, float *, int64_t, int64_t *) { return fail("not implemented yet"); }
int wdb_topk(int, void *, const wdb_col_t *, int, const char *, const char *, const char *, int, int64_t, int64_t, int64_t, float *, float *, int64_t *) { return fail("not implemented yet"); }
int wdb_sort_float(int, void *, float *, int64_t, int) { return fail("not implemented yet"); }
int wdb_sort_pairs(int, void *, int32_t *, float *, int64_t, int) { return fail("not implemented yet"); }
}
