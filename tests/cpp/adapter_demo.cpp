// Drives the C++ adapters (include/icp_b200_engine.hpp) the way the reference's callers drive the reference:
//   engine mode : RegistrationService::startRegistration (services/registrationservice.cpp:186-213) -> ICPEngine
//   cli mode    : main() of icp_registration.cpp (:912) -> ICP(...)
// usage: adapter_demo <engine|cli> <src.bin> <tgt.bin> <max_iterations> <tolerance>     (bin = raw double xyz triples)
// Prints one line of JSON; the moved source is written to <src.bin>.out
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "icp_b200_engine.hpp"

struct Point3D {  // core/pointcloud.h:12-23
    double x, y, z;
};
struct PointCloud {  // the part of core/pointcloud.h:30-65 the engine touches
    std::vector<Point3D> points;
};

static PointCloud load(const char* path) {
    PointCloud c;
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return c;
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    c.points.resize((size_t)bytes / sizeof(Point3D));
    f.read(reinterpret_cast<char*>(c.points.data()), bytes);
    return c;
}

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const std::string mode = argv[1];
    PointCloud src = load(argv[2]), tgt = load(argv[3]);
    const int iters = std::atoi(argv[4]);
    const double tol = std::atof(argv[5]);
    std::string json = "{";
    if (mode == "engine") {
        icpb200::ICPEngine eng;
        icpb200::ICPParameters p;
        p.maxIterations = iters;
        p.tolerance = tol;
        eng.setParameters(p);
        std::string order, fin_msg;
        bool fin_ok = false;
        eng.started = [&] { order += "s"; };
        eng.iterationCompleted = [&](const icpb200::IterationResult&) { order += "i"; };
        eng.progressUpdated = [&](int, int, double) { order += "p"; };
        eng.finished = [&](bool ok, const std::string& m) { fin_ok = ok; fin_msg = m; order += "f"; };
        eng.registerPointClouds(&src, &tgt);
        const icpb200::ICPResult r = eng.getResult();
        char buf[512];
        std::snprintf(buf, sizeof buf, "\"success\": %s, \"finished_ok\": %s, \"totalIterations\": %d, \"finalRMSE\": %.17g, \"order\": \"%s\", ",
                      r.success ? "true" : "false", fin_ok ? "true" : "false", r.totalIterations, r.finalRMSE, order.c_str());
        json += buf;
        json += "\"finalR\": [";
        for (int i = 0; i < 9; ++i) {
            std::snprintf(buf, sizeof buf, "%s%.17g", i ? ", " : "", r.finalR[i / 3][i % 3]);
            json += buf;
        }
        json += "], \"finalT\": [";
        for (int i = 0; i < 3; ++i) {
            std::snprintf(buf, sizeof buf, "%s%.17g", i ? ", " : "", r.finalT[i]);
            json += buf;
        }
        json += "]";
    } else {
        double R[3][3], t[3];
        std::vector<icpb200::Mat4> its;
        icpb200::ICP(src, tgt, iters, tol, R, t, &its);
        char buf[256];
        std::snprintf(buf, sizeof buf, "\"n_transforms\": %zu, \"finalR\": [", its.size());
        json += buf;
        for (int i = 0; i < 9; ++i) {
            std::snprintf(buf, sizeof buf, "%s%.17g", i ? ", " : "", R[i / 3][i % 3]);
            json += buf;
        }
        json += "], \"finalT\": [";
        for (int i = 0; i < 3; ++i) {
            std::snprintf(buf, sizeof buf, "%s%.17g", i ? ", " : "", t[i]);
            json += buf;
        }
        json += "]";
    }
    json += "}";
    std::puts(json.c_str());
    std::ofstream o(std::string(argv[2]) + ".out", std::ios::binary);
    o.write(reinterpret_cast<const char*>(src.points.data()), (std::streamsize)(src.points.size() * sizeof(Point3D)));
    return 0;
}
