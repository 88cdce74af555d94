// The reference's call sequence for one overlap pair (CPU/main.cpp:82-89), written against include/pixflow_b200.hpp
// exactly as it is written against the reference's OpticalFlow.hpp: new NovelViewGeneratorAsymmetricFlow(name) ->
// prepare(L, R) -> setBlend(blend) -> generateNovelView(out) -> delete.  Inputs/outputs are raw files so that the
// Python test can check the results against the CPU oracle.
//
//   novel_view_main <flow_alg> <rows> <cols> <L.bgra> <R.bgra> <blend.f32> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "pixflow_b200.hpp"

using namespace optical_flow;

static bool read_file(const std::string& path, void* dst, size_t bytes) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    const size_t n = fread(dst, 1, bytes, f);
    fclose(f);
    return n == bytes;
}
static bool write_file(const std::string& path, const Mat& m, size_t elem) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    for (int y = 0; y < m.rows; ++y) fwrite(m.data + (size_t)y * m.step, 1, (size_t)m.cols * elem, f);
    fclose(f);
    return true;
}

int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage: %s alg rows cols L R blend out_prefix\n", argv[0]); return 2; }
    const std::string alg = argv[1];
    const int rows = atoi(argv[2]), cols = atoi(argv[3]);
    Mat overlappedL = pf::make_mat(rows, cols, pf::PF_8UC4), overlappedR = pf::make_mat(rows, cols, pf::PF_8UC4);
    Mat blend = pf::make_mat(rows, cols, pf::PF_32FC1);
    if (!read_file(argv[4], overlappedL.data, (size_t)rows * cols * 4) || !read_file(argv[5], overlappedR.data, (size_t)rows * cols * 4) ||
        !read_file(argv[6], blend.data, (size_t)rows * cols * 4)) { fprintf(stderr, "cannot read inputs\n"); return 2; }
    try {
        NovelViewGenerator* novelViewGen = new NovelViewGeneratorAsymmetricFlow(alg);
        novelViewGen->prepare(overlappedL, overlappedR);
        novelViewGen->setBlend(blend);
        Mat novelViewMerged = Mat();
        novelViewGen->generateNovelView(novelViewMerged);
        const std::string out = argv[7];
        write_file(out + ".flowLR", novelViewGen->getFlowLtoR(), 8);
        write_file(out + ".flowRL", novelViewGen->getFlowRtoL(), 8);
        write_file(out + ".merged", novelViewMerged, 4);
        delete novelViewGen;
        // the factory + abstract interface, CPU/OpticalFlow.cpp:128-141
        OpticalFlowInterface* flowAlg = makeOpticalFlowByName(alg);
        Mat flow;
        flowAlg->computeOpticalFlow(overlappedL, overlappedR, flow, OpticalFlowInterface::DirectionHint::LEFT);
        write_file(out + ".flow", flow, 8);
        delete flowAlg;
    } catch (const util::VrCamException& e) {
        fprintf(stderr, "VrCamException: %s\n", e.what());
        return 3;
    }
    return 0;
}
