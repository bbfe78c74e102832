// makedb - drop-in for the reference's database builder (reference src/makedb.cpp:279-374): FASTA/FASTQ(.gz) -> the
// six binary files + two metadata files the scan engine (ours and the reference's) reads. Output is byte-identical to
// the reference's for the same input: same padding (to 4, code 20), same letter map (src/convert.cuh:6-34), same
// std::sort-by-length call on the same index vector (src/makedb.cpp:185-195), same metadata layout (:214-222).
//   usage: makedb input.fa(.gz) dbprefix [options ignored: --mem val, --tempdir dir]
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

#include "fasta_reader.hpp"

static const int kBoundaries[36] = {48,  64,  80,  96,  112, 128, 144, 160, 176, 192,  208,  224,
                                    240, 256, 288, 320, 352, 384, 416, 448, 480, 512,  576,  640,
                                    704, 768, 832, 896, 960, 1024, 1088, 1152, 1216, 1280, 8000, 2147483646};

static char encodeResidue(char c) {
    static const char order[] = "ARNDCQEGHILKMFPSTWYV";
    for (int i = 0; i < 20; i++)
        if (c == order[i]) return (char)i;
    return 20;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::cout << "Usage: " << argv[0] << " pathtodb(fasta/fastq, gzip ok) outputprefix [--mem val] [--tempdir dir]\n";
        return 0;
    }
    const std::string input = argv[1], prefix = argv[2];
    std::vector<char> chars, headers;
    std::vector<size_t> offsets{0}, headerOffsets{0};
    std::vector<int32_t> lengths;
    try {
        sw4::SequenceFileReader reader(input);
        while (reader.next()) {
            const std::string& s = reader.getCurrentSequence();
            const std::string& h = reader.getCurrentHeader();
            for (char c : s) chars.push_back(encodeResidue(c));
            const size_t pad = (s.size() % 4 == 0) ? 0 : 4 - s.size() % 4;
            chars.insert(chars.end(), pad, (char)20);
            offsets.push_back(chars.size());
            lengths.push_back((int32_t)s.size());
            headers.insert(headers.end(), h.begin(), h.end());
            headerOffsets.push_back(headers.size());
        }
    } catch (const std::exception& e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
    const size_t n = lengths.size();
    std::vector<int32_t> indices(n);
    std::iota(indices.begin(), indices.end(), 0);
    std::sort(indices.begin(), indices.end(), [&](const auto& l, const auto& r) { return lengths[l] < lengths[r]; });

    std::vector<size_t> perPartition(36, 0);
    {
        size_t pos = 0;
        for (int p = 0; p < 36; p++) {
            size_t end = pos;
            while (end < n && lengths[indices[end]] <= kBoundaries[p]) end++;
            perPartition[p] = end - pos;
            pos = end;
        }
    }
    auto open = [&](const std::string& name) {
        std::ofstream f(prefix + name, std::ios::binary);
        if (!f) { std::cerr << "Cannot open output file " << prefix + name << "\n"; exit(1); }
        return f;
    };
    {
        auto meta = open("0metadata");
        const int np = 36;
        meta.write((const char*)&np, sizeof(int));
        meta.write((const char*)kBoundaries, sizeof(int) * 36);
        meta.write((const char*)perPartition.data(), sizeof(size_t) * 36);
    }
    auto fh = open("0headers"), fho = open("0headeroffsets"), fc = open("0chars"), fo = open("0offsets"), fl = open("0lengths");
    size_t curH = 0, curC = 0;
    fho.write((const char*)&curH, sizeof(size_t));
    fo.write((const char*)&curC, sizeof(size_t));
    for (size_t i = 0; i < n; i++) {
        const size_t s = (size_t)indices[i];
        const size_t hl = headerOffsets[s + 1] - headerOffsets[s];
        fh.write(headers.data() + headerOffsets[s], (std::streamsize)hl);
        curH += hl;
        fho.write((const char*)&curH, sizeof(size_t));
        const size_t nc = offsets[s + 1] - offsets[s];
        fc.write(chars.data() + offsets[s], (std::streamsize)nc);
        fl.write((const char*)&lengths[s], sizeof(int32_t));
        curC += nc;
        fo.write((const char*)&curC, sizeof(size_t));
    }
    open("metadata");  // empty global metadata file (src/dbdata.cpp:192-197)
    std::cout << "Parsed " << n << " sequences\n";
    return 0;
}
