// The loop body of the reference driver (CPU/main.cpp:72-95) written against include/pixflow_b200.hpp exactly as it is written
// against the reference's StitchTool.hpp + OpticalFlow.hpp: Stitchtools::prepare -> getOverlappedL/R, getBlend ->
// NovelViewGeneratorAsymmetricFlow::prepare -> setBlend -> generateNovelView -> setMergedmiddle -> Gather -> getFinalResult,
// then the same iteration through the fused device-resident call.  Raw files in and out for the Python test.
//
//   stitch_main <flow_alg> <rows> <cols> <L.bgra> <R.bgra> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "pixflow_b200.hpp"

using namespace optical_flow;
using namespace stitch_tools;

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
    if (argc != 7) { fprintf(stderr, "usage: %s alg rows cols L R out_prefix\n", argv[0]); return 2; }
    const std::string FLAGS_flow_alg = argv[1];
    const int rows = atoi(argv[2]), cols = atoi(argv[3]);
    Mat colorImageL = pf::make_mat(rows, cols, pf::PF_8UC4), colorImageR = pf::make_mat(rows, cols, pf::PF_8UC4);
    if (!read_file(argv[4], colorImageL.data, (size_t)rows * cols * 4) || !read_file(argv[5], colorImageR.data, (size_t)rows * cols * 4)) {
        fprintf(stderr, "cannot read inputs\n");
        return 2;
    }
    const std::string out = argv[6];
    try {
        Stitchtools Stools;
        Stools.prepare(colorImageL, colorImageR);

        Mat overlappedL = Stools.getOverlappedL();
        Mat overlappedR = Stools.getOverlappedR();
        Mat blend = Stools.getBlend();

        NovelViewGenerator* novelViewGen = new NovelViewGeneratorAsymmetricFlow(FLAGS_flow_alg);
        novelViewGen->prepare(overlappedL, overlappedR);

        novelViewGen->setBlend(blend);
        Mat novelViewMerged = Mat();
        novelViewGen->generateNovelView(novelViewMerged);

        Stools.setMergedmiddle(novelViewMerged);
        Stools.Gather();
        Mat FinalResult = Stools.getFinalResult();
        delete novelViewGen;

        write_file(out + ".map", Stools.getMap(), 1);
        write_file(out + ".blend", blend, 4);
        write_file(out + ".final", FinalResult, 4);

        PixFlowB200 flowAlg(FLAGS_flow_alg);
        Mat fused;
        stitchIteration(flowAlg, colorImageL, colorImageR, fused);
        write_file(out + ".fused", fused, 4);
    } catch (const util::VrCamException& e) {
        fprintf(stderr, "VrCamException: %s\n", e.what());
        return 3;
    }
    return 0;
}
