// SURVEY.md §8 row f4 — Swin sparse-token grouping off the Python critical path (host code; no device work).
//
// The reference packs the visible tokens of the local windows into groups of at most group_size tokens with a 0/1
// knapsack run greedily until no window is left (GreenMIM-style), in pure Python, once per forward and BasicBlock:
//   knapsack(W, wt)                    model/sub_module/swin_block.py:280-326
//   group_windows(group_size, wt)      model/sub_module/swin_block.py:329-352
// This restates both with the same table, the same back-tracking rule and the same tie behaviour, so the grouping is
// identical window for window.  (The mask is batch-shared, swin.py:158: the Python side caches the plan per mask.)
#include <vector>

#include "ep_common.cuh"

namespace ep {
namespace {

// returns the best fill and appends the selected item positions (increasing) to `sel`
int knapsack(int W, const std::vector<int>& wt, std::vector<int>& sel) {
    const int n = (int)wt.size();
    std::vector<int> K((size_t)(n + 1) * (W + 1), 0);
    auto at = [&](int i, int w) -> int& { return K[(size_t)i * (W + 1) + w]; };
    for (int i = 1; i <= n; ++i)
        for (int w = 1; w <= W; ++w) {
            int best = at(i - 1, w);
            if (wt[i - 1] <= w) {
                const int take = wt[i - 1] + at(i - 1, w - wt[i - 1]);
                if (take > best) best = take;      // max(take, skip): equal values keep either, the table entry is the same
            }
            at(i, w) = best;
        }
    const int res_ret = at(n, W);
    int res = res_ret, w = W;
    std::vector<int> rev;
    for (int i = n; i > 0; --i) {
        if (res <= 0) break;
        if (res == at(i - 1, w)) continue;         // the value came from the row above: item i-1 not included
        rev.push_back(i - 1);
        res -= wt[i - 1];
        w -= wt[i - 1];
    }
    sel.assign(rev.rbegin(), rev.rend());
    return res_ret;
}

}  // namespace
}  // namespace ep

extern "C" {

int ep_swin_group_windows_host(int group_size, const int* num_ele_win, int n_win, int* num_ele_group, int* group_first,
                               int* grouped_idx, int* n_groups) {
    if (group_size <= 0 || n_win < 0 || (n_win > 0 && (!num_ele_win || !num_ele_group || !group_first || !grouped_idx)) || !n_groups)
        return EP_EINVAL;
    for (int i = 0; i < n_win; ++i)
        if (num_ele_win[i] <= 0 || num_ele_win[i] > group_size) return EP_EINVAL;     // the reference loops forever on such input
    std::vector<int> wt(num_ele_win, num_ele_win + n_win), ori(n_win);
    for (int i = 0; i < n_win; ++i) ori[i] = i;
    int ng = 0, filled = 0;
    if (n_win > 0) group_first[0] = 0;
    while (!wt.empty()) {
        std::vector<int> sel;
        const int res = ep::knapsack(group_size, wt, sel);
        num_ele_group[ng] = res;
        for (int s : sel) grouped_idx[filled++] = ori[s];
        group_first[++ng] = filled;
        // drop the selected windows, keep the order of the rest
        std::vector<int> wt2, ori2;
        size_t k = 0;
        for (int i = 0; i < (int)wt.size(); ++i) {
            if (k < sel.size() && sel[k] == i) { ++k; continue; }
            wt2.push_back(wt[i]);
            ori2.push_back(ori[i]);
        }
        wt.swap(wt2);
        ori.swap(ori2);
    }
    *n_groups = ng;
    return EP_OK;
}

int ep_swin_group_tables_host(const int64_t* group_id, const int64_t* coords, int n_groups, int group_size, int window, int mask_rel,
                              float* attn_mask, int64_t* rel_pos_idx) {
    if (!group_id || !coords || !attn_mask || !rel_pos_idx || n_groups <= 0 || group_size <= 0 || window <= 0) return EP_EINVAL;
    const int64_t span = 2 * (int64_t)window - 1;
    for (int g = 0; g < n_groups; ++g) {
        const int64_t* gid = group_id + (size_t)g * group_size;
        const int64_t* c = coords + (size_t)g * group_size * 2;
        float* am = attn_mask + (size_t)g * group_size * group_size;
        int64_t* rp = rel_pos_idx + (size_t)g * group_size * group_size;
        for (int i = 0; i < group_size; ++i) {
            const float gi = (float)gid[i];                       // the reference compares the ids as float32
            for (int j = 0; j < group_size; ++j) {
                // two slots attend to each other when they carry the same window id and are not both padding (id -1)
                const bool blocked = (gi - (float)gid[j]) != 0.0f || (gid[i] == -1 && gid[j] == -1);
                am[(size_t)i * group_size + j] = blocked ? -100.0f : 0.0f;
                const int64_t rel = (c[2 * i] - c[2 * j] + window - 1) * span + (c[2 * i + 1] - c[2 * j + 1] + window - 1);
                rp[(size_t)i * group_size + j] = (mask_rel && blocked) ? 0 : rel;
            }
        }
    }
    return EP_OK;
}


}  // extern "C"
