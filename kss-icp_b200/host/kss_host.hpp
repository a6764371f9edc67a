// kss_host.hpp -- glue between the reference-shaped C++ classes and the C ABI (include/kss_icp_b200.h).
#pragma once
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "kss_icp_b200.h"

namespace kss_host {

typedef std::vector<std::vector<double>> Cloud;   // the reference's cloud type (KSS_ICP.hpp:42-49)

// one context per process (device from KSS_DEVICE, default 0); the reference is single threaded here
inline kss_ctx* ctx() {
    static kss_ctx* c = nullptr;
    if (!c) {
        const char* d = std::getenv("KSS_DEVICE");
        int rc = kss_ctx_create(d ? std::atoi(d) : 0, &c);
        if (rc != KSS_OK) {
            std::fprintf(stderr, "[kss-icp_b200] kss_ctx_create failed (%d): a CUDA device is required, there is no CPU path\n", rc);
            std::abort();
        }
    }
    return c;
}

inline std::vector<double> pack(const Cloud& c) {
    std::vector<double> f(c.size() * 3);
    for (size_t i = 0; i < c.size(); ++i) { f[3 * i] = c[i][0]; f[3 * i + 1] = c[i][1]; f[3 * i + 2] = c[i][2]; }
    return f;
}
inline Cloud unpack(const std::vector<double>& f) {
    Cloud c(f.size() / 3, std::vector<double>(3));
    for (size_t i = 0; i < c.size(); ++i) { c[i][0] = f[3 * i]; c[i][1] = f[3 * i + 1]; c[i][2] = f[3 * i + 2]; }
    return c;
}
// the reference never checks errors: report on stderr and keep its behaviour
inline bool ok(int rc, const char* what) {
    if (rc == KSS_OK) return true;
    std::fprintf(stderr, "[kss-icp_b200] %s failed (%d): %s\n", what, rc, kss_last_error(ctx()));
    return false;
}

}  // namespace kss_host
