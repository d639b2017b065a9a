// TEST INFRASTRUCTURE: host stand-in for the two CUB entry points the backend uses (see ../../cuda_runtime.h)
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>
namespace cub {
struct DeviceRadixSort {
    template <class K> static int SortKeys(void *tmp, size_t &bytes, const K *in, K *out, int n, int = 0, int = sizeof(K) * 8, void * = nullptr) {
        if (!tmp) { bytes = 16; return 0; }
        std::vector<K> v(in, in + n);
        std::stable_sort(v.begin(), v.end());
        std::copy(v.begin(), v.end(), out);
        return 0;
    }
    template <class K, class V> static int SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int = 0, int = sizeof(K) * 8, void * = nullptr) {
        if (!tmp) { bytes = 16; return 0; }
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return kin[a] < kin[b]; });
        std::vector<K> k(n); std::vector<V> v(n);
        for (int i = 0; i < n; ++i) { k[i] = kin[idx[i]]; v[i] = vin[idx[i]]; }
        std::copy(k.begin(), k.end(), kout); std::copy(v.begin(), v.end(), vout);
        return 0;
    }
};
} // namespace cub
