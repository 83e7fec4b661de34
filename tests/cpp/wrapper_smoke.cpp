/* Compiles against include/pheniqs_b200.hpp and exercises the host-only half of the C ABI
   (configuration + packing) through the C++ wrapper; used by tests/test_cpp_wrapper.py. */
#include <pheniqs_b200.hpp>

#include <cstdio>
#include <cstring>

int main(int argc, char** argv) {
    const char* job =
        "{\"sample\": {\"algorithm\": \"mdd\", \"transform\": {\"token\": [\"0::4\"]},"
        " \"codec\": {\"@b\": {\"barcode\": [\"ACGT\"]}, \"@a\": {\"barcode\": [\"TTTT\"]}}}}";
    try {
        const std::string compiled(phq::compile_job(job));
        if(compiled.find("\"distance tolerance\"") == std::string::npos) { std::printf("no tolerance\n"); return 1; }
        /* device -1: host-only handle */
        phq::BatchDecoder decoder(compiled, -1);
        if(decoder.decoder_cardinality() != 1 || decoder.info(0).barcode_cardinality != 2 || decoder.info(0).nucleotide_cardinality != 4) { return 2; }
        phq::TileBuffer buffer;
        buffer.allocate(decoder.info(0), 2, false);
        const uint8_t code[8] = { 1, 2, 4, 8, 8, 15, 8, 8 };
        const uint8_t quality[8] = { 30, 30, 30, 30, 10, 2, 10, 10 };
        const int64_t offset[3] = { 0, 4, 8 };
        const uint8_t* pc[1] = { code };
        const uint8_t* pq[1] = { quality };
        const int64_t* po[1] = { offset };
        std::vector< phq_tile > tiles(1, buffer.tile());
        decoder.pack(2, 1, pc, pq, po, tiles);
        /* read 0 = ACGT: lo plane 0b1010 (C, T), hi plane 0b1100 (G, T); read 1 has an N at position 1 */
        if(tiles[0].bases[0] != (0xau | (0xcu << 16)) || tiles[0].nmask[0] != 0 || tiles[0].nmask[1] != 2) { return 3; }
        if(tiles[0].quality[0] != 0x1e1e1e1eu || tiles[0].quality_bits != 8) { return 4; }
        /* the same batch with the smallest quality form: three distinct values (30, 10, 2) -> 2-bit indices */
        tiles[0].quality_bits = -1;
        decoder.pack(2, 1, pc, pq, po, tiles);
        if(tiles[0].quality_bits != 2 || tiles[0].quality_codebook[0] != 2 || tiles[0].quality_codebook[1] != 10 || tiles[0].quality_codebook[2] != 30) { return 7; }
        if(tiles[0].quality[0] != 0xaau || tiles[0].quality[1] != 0x51u) { return 8; }
        bool refused(false);
        try {
            std::vector< phq_result > results(2);
            std::vector< phq_result* > pr(1, results.data());
            uint8_t qc[2];
            decoder.classify(2, tiles, NULL, pr, qc);
        } catch(const phq::InternalError&) { refused = true; }       /* no CPU classification path */
        if(!refused) { return 5; }
        bool rejected(false);
        try { phq::compile_job("{\"sample\": {\"algorithm\": \"mdd\", \"transform\": {\"token\": [\"0::4\"]}, \"distance tolerance\": [3], \"codec\": {\"@a\": {\"barcode\": [\"ACGT\"]}, \"@b\": {\"barcode\": [\"ACGA\"]}}}}"); }
        catch(const phq::ConfigurationError& e) { rejected = e.code == 3; }
        if(!rejected) { return 6; }
        /* a job file of the reference's own test (import + base inheritance), when one is given */
        if(argc > 1) {
            const std::string loaded(phq::load_job(argv[1]));
            if(loaded.find("\"import\"") != std::string::npos || loaded.find("BDGGG_sample") == std::string::npos) { return 9; }
            const std::string decoders(phq::compile_job(loaded));
            if(decoders.find("\"ID\": \"BDGGG:1:AGGCAGAA\"") == std::string::npos && decoders.find("\"ID\":\"BDGGG:1:AGGCAGAA\"") == std::string::npos) { return 11; }
            phq::BatchDecoder chain(decoders, -1);
            if(chain.decoder_cardinality() != 3) { return 12; }
            /* RG 24 + BC / QT 24 + XB 7, OX / BZ 24, CB 12 + CR / CY 24 + XC 7 = 122 bytes, in strides of 16 */
            if(chain.tag_record_bytes() != 128) { return 13; }
        }
    } catch(const phq::Error& e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 10;
    }
    std::printf("ok\n");
    return 0;
}
