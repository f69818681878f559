// Host-only helpers of the setup phase (no device work): sequential sweeps that are O(n) but
// cost seconds as Python loops at the configured sizes.
#include "common.cuh"

// Phase 1 of the reference's aggregation (PySolvers/Linear/SmoothedAggregation.py:72-89), after the
// isolated nodes (|N_i| = 1) have founded their aggregates in index order: sweep i = 0 .. n-1; a
// free node whose whole strong neighbourhood N_i (CSR lists s_ptr / s_cols, i itself included
// implicitly) is still free founds the next aggregate and takes N_i with it.  `agg_of` comes in
// with the isolated nodes already assigned (-1 = free), `n_agg` with their count; roots of the new
// aggregates are appended to h_roots (capacity n).  Returns the new number of aggregates via *n_agg.
extern "C" int psb_sa_phase1(int64_t n, const int64_t* h_s_ptr, const int64_t* h_s_cols, int64_t* h_agg_of,
                             int64_t* h_roots, int64_t* n_agg) {
  PSB_REQUIRE(n >= 0 && h_s_ptr && h_agg_of && h_roots && n_agg && (h_s_ptr[n] == 0 || h_s_cols), PSB_ERR_ARG,
              "psb_sa_phase1: NULL argument");
  int64_t count = *n_agg;
  PSB_REQUIRE(count >= 0 && count <= n, PSB_ERR_ARG, "psb_sa_phase1: bad aggregate count");
  for (int64_t i = 0; i < n; ++i) {
    if (h_agg_of[i] >= 0) continue;
    const int64_t a = h_s_ptr[i], b = h_s_ptr[i + 1];
    bool ok = true;
    for (int64_t k = a; k < b; ++k)
      if (h_agg_of[h_s_cols[k]] >= 0) { ok = false; break; }
    if (!ok) continue;
    for (int64_t k = a; k < b; ++k) h_agg_of[h_s_cols[k]] = count;
    h_agg_of[i] = count;
    h_roots[count] = i;
    ++count;
  }
  *n_agg = count;
  return PSB_OK;
}
