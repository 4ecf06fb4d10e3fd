/* oracle/link_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, never on the product path).
 *
 * Linking of localizations into binding events, picasso.postprocess (reference
 * picasso/postprocess.py @ 96e0da51): `_get_link_groups` :2440-2507 with
 * `_get_next_loc_index_in_link_group` :2510-2552 (greedy, frame-ordered: a chain takes the first
 * not yet linked localization of the same `group` within d_max in the next max_dark_time + 1
 * frames) and the per-group loops `_link_group_count` :2555, `_link_group_sum` :2567,
 * `_link_group_min_max` :2618, `_link_group_last` :2648 (sequential accumulation in the column's
 * dtype, localization order).  numba types: coordinates keep their dtype (float32 or float64),
 * squares are formed in that dtype and compared with the float64 d_max**2.
 * Pinned by tests/golden/link.npz.
 */
#include <stddef.h>

#define NEXT_IN_GROUP(T, NAME)                                                                        \
    static long long NAME(long long cur, const int* link_group, long long N, const long long* frame,    \
                          const T* x, const T* y, double d_max, long long max_dark, const int* group) { \
        const long long cf = frame[cur];                                                                \
        const T cx = x[cur], cy = y[cur];                                                               \
        const int cg = group[cur];                                                                      \
        const long long min_frame = cf + 1;                                                             \
        long long min_index = cur + 1;                                                                  \
        /* numba: `for min_index in range(cur + 1, N): if frame >= min_frame: break` leaves the last   \
           value N - 1 when nothing breaks (or cur + 1 .. when the range is empty: stays undefined ->  \
           the reference then reads the stale variable; with cur == N - 1 the range is empty and       \
           min_index keeps its previous value in numba's lowering = 0-initialised slot) */             \
        int found = 0;                                                                                  \
        for (long long k = cur + 1; k < N; k++) { min_index = k; if (frame[k] >= min_frame) { found = 1; break; } } \
        if (cur + 1 >= N) return -1;                                                                    \
        (void)found;                                                                                    \
        const long long max_frame = cf + max_dark + 1;                                                  \
        long long max_index = N;                                                                        \
        for (long long k = min_index; k < N; k++) if (frame[k] > max_frame) { max_index = k; break; }   \
        const double d2 = d_max * d_max;                                                                \
        for (long long j = min_index; j < max_index; j++) {                                             \
            if (group[j] != cg || link_group[j] != -1) continue;                                        \
            const T dx = cx - x[j], dy = cy - y[j];                                                     \
            const T dx2 = dx * dx;                                                                      \
            if (!((double)dx2 <= d2)) continue;                                                         \
            const T dy2 = dy * dy;                                                                      \
            if (!((double)dy2 <= d2)) continue;                                                         \
            const T s = dx2 + dy2;                                                                      \
            if ((double)s <= d2) return j;                                                              \
        }                                                                                               \
        return -1;                                                                                      \
    }

NEXT_IN_GROUP(float, next_f32)
NEXT_IN_GROUP(double, next_f64)

int orc_get_link_groups(long long N, const long long* frame, const void* x, const void* y, int f64,
                        double d_max, long long max_dark, const int* group, int* link_group) {
    for (long long i = 0; i < N; i++) link_group[i] = -1;
    int current = -1;
    for (long long i = 0; i < N; i++) {
        if (link_group[i] != -1) continue;
        current++;
        link_group[i] = current;
        long long cur = i;
        for (;;) {
            const long long nxt = f64 ? next_f64(cur, link_group, N, frame, (const double*)x, (const double*)y,
                                                 d_max, max_dark, group)
                                      : next_f32(cur, link_group, N, frame, (const float*)x, (const float*)y,
                                                 d_max, max_dark, group);
            if (nxt == -1) break;
            link_group[nxt] = current;
            cur = nxt;
        }
    }
    return current + 1;
}

/* sequential per-group accumulation in localization order (dtype: 0 f32, 1 f64, 2 u32, 3 i32) */
int orc_link_group_sum(long long N, const int* link_group, int n_groups, int dtype, const void* col, void* out) {
    (void)n_groups;
    for (long long i = 0; i < N; i++) {
        const int g = link_group[i];
        switch (dtype) {
            case 0: ((float*)out)[g] += ((const float*)col)[i]; break;
            case 1: ((double*)out)[g] += ((const double*)col)[i]; break;
            case 2: ((unsigned*)out)[g] += ((const unsigned*)col)[i]; break;
            default: ((int*)out)[g] += ((const int*)col)[i]; break;
        }
    }
    return 0;
}
