// Minimal FASTA / FASTQ reader with transparent gzip (zlib), for the drop-in CLIs. It reproduces what the reference's
// kseqpp wrapper hands to its callers (src/kseqpp/kseqpp.hpp:54-128): header = the whole line after '>' or '@',
// sequence = concatenated lines, quality ignored.
#pragma once
#include <zlib.h>

#include <stdexcept>
#include <string>

namespace sw4 {

class SequenceFileReader {
public:
    explicit SequenceFileReader(const std::string& path) {
        file = gzopen(path.c_str(), "rb");
        if (!file) throw std::runtime_error("Cannot open file " + path);
        gzbuffer(file, 1 << 20);
    }
    ~SequenceFileReader() { if (file) gzclose(file); }
    SequenceFileReader(const SequenceFileReader&) = delete;
    SequenceFileReader& operator=(const SequenceFileReader&) = delete;

    // returns false at end of file
    bool next() {
        header.clear(); sequence.clear();
        if (!havePending) { if (!readLine(pending)) return false; }
        havePending = false;
        while (pending.empty()) { if (!readLine(pending)) return false; }
        if (pending[0] != '>' && pending[0] != '@') throw std::runtime_error("unexpected line in sequence file: " + pending);
        const bool fastq = pending[0] == '@';
        header = pending.substr(1);
        std::string line;
        if (!fastq) {
            while (readLine(line)) {
                if (!line.empty() && line[0] == '>') { pending = line; havePending = true; break; }
                sequence += line;
            }
        } else {
            while (readLine(line)) {
                if (!line.empty() && line[0] == '+') break;
                sequence += line;
            }
            size_t q = 0;
            while (q < sequence.size() && readLine(line)) q += line.size();
        }
        return true;
    }
    const std::string& getCurrentHeader() const { return header; }
    const std::string& getCurrentSequence() const { return sequence; }

private:
    bool readLine(std::string& out) {
        out.clear();
        char buf[1 << 16];
        bool any = false;
        while (gzgets(file, buf, sizeof(buf))) {
            any = true;
            out += buf;
            if (!out.empty() && out.back() == '\n') break;
        }
        while (!out.empty() && (out.back() == '\n' || out.back() == '\r')) out.pop_back();
        return any;
    }
    gzFile file = nullptr;
    std::string header, sequence, pending;
    bool havePending = false;
};

}  // namespace sw4
