/* The C++ wrapper (include/pheniqs_b200.hpp) classifying on a GPU: a C++ host with no Python in the loop.
   Used by tests/test_gpu_binding.py. Two barcodes ACGT / TTTT, MDD and PAMLD; four reads with known verdicts. */
#include <pheniqs_b200.hpp>

#include <cmath>
#include <cstdio>

static int run(const char* algorithm) {
    const std::string job(std::string("{\"sample\": {\"algorithm\": \"") + algorithm + "\", \"transform\": {\"token\": [\"0::4\"]},"
        " \"codec\": {\"@b\": {\"barcode\": [\"ACGT\"]}, \"@a\": {\"barcode\": [\"TTTT\"]}}}}");
    const std::string compiled(phq::compile_job(job));
    phq::BatchDecoder decoder(compiled, 0);
    phq::TileBuffer buffer;
    buffer.allocate(decoder.info(0), 4, true);
    /* TTTT (index 1: key @a sorts first), ACGT (index 2), ACGA (one mismatch from ACGT), NNNN */
    const uint8_t code[16] = { 8, 8, 8, 8,  1, 2, 4, 8,  1, 2, 4, 1,  15, 15, 15, 15 };
    const uint8_t quality[16] = { 37, 37, 37, 37,  37, 37, 37, 37,  37, 37, 37, 12,  2, 2, 2, 2 };
    const int64_t offset[5] = { 0, 4, 8, 12, 16 };
    const uint8_t* pc[1] = { code };
    const uint8_t* pq[1] = { quality };
    const int64_t* po[1] = { offset };
    std::vector< phq_tile > tiles(1, buffer.tile());
    decoder.pack(4, 1, pc, pq, po, tiles);
    std::vector< phq_result > results(4);
    std::vector< phq_result* > pr(1, results.data());
    uint8_t qc[4];
    decoder.classify(4, tiles, NULL, pr, qc);
    if(results[0].index != 1 || results[1].index != 2 || results[0].distance != 0 || results[1].distance != 0) { return 1; }
    const bool pamld(std::string(algorithm) == "pamld");
    if(pamld) {
        if(!(results[0].confidence > 0.99 && results[0].confidence < 1.0)) { return 2; }
        if(results[2].index != 2 || results[2].distance != 1) { return 3; }
        if(results[3].index != 0 || qc[3] != 1) { return 4; }
    } else {
        if(results[0].confidence != 0.0) { return 2; }
        if(results[2].index != 2 || results[2].distance != 1 || qc[2] != 0) { return 3; }       /* tolerance 1 = the Shannon bound of a distance 3 pair */
        if(results[3].index != 0 || qc[3] != 1) { return 4; }
    }
    std::vector< uint64_t > u;
    std::vector< double > f;
    decoder.accumulators(0, u, f);
    uint64_t total(0);
    for(size_t row(0); row < 3; ++row) { total += u[row * 6]; }
    if(total != 4) { return 5; }
    uint64_t count(0), pf(0);
    decoder.totals(count, pf);
    if(count != 4) { return 6; }
    const std::string report(decoder.report(4, 4));
    if(report.find("\"sample\"") == std::string::npos) { return 7; }
    /* the same reads as the reference keeps them (BAM codes + Phred bytes), packed on the device */
    std::vector< phq_raw_segment > segment(1);
    segment[0].sequence = code; segment[0].quality = quality; segment[0].offset = offset; segment[0].length = 0;
    std::vector< phq_result > again(4);
    std::vector< phq_result* > pa(1, again.data());
    uint8_t qc_again[4];
    decoder.classify_bam(4, segment, NULL, pa, qc_again);
    for(int r(0); r < 4; ++r) {
        if(again[r].index != results[r].index || again[r].distance != results[r].distance || again[r].confidence != results[r].confidence || qc_again[r] != qc[r]) { return 8; }
    }
    return 0;
}

int main() {
    try {
        const int a(run("mdd"));
        if(a != 0) { std::printf("mdd failed at %d\n", a); return a; }
        const int b(run("pamld"));
        if(b != 0) { std::printf("pamld failed at %d\n", b); return 10 + b; }
    } catch(const phq::Error& e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 100;
    }
    std::printf("ok\n");
    return 0;
}
