// TEST INFRASTRUCTURE: host stand-in for cub::DeviceSelect::Unique
#pragma once
#include <cstddef>
namespace cub {
struct DeviceSelect {
    template <class T, class C> static int Unique(void *tmp, size_t &bytes, const T *in, T *out, C *num_out, int n, void * = nullptr) {
        if (!tmp) { bytes = 16; return 0; }
        int m = 0;
        for (int i = 0; i < n; ++i) if (i == 0 || !(in[i] == in[i - 1])) out[m++] = in[i];
        *num_out = (C)m;
        return 0;
    }
};
} // namespace cub
