// inputfile_check.cc -- test driver: streams argv[1] through goss::InputFile (plain, .gz, .bz2, "-") in odd-sized reads
// to stdout; exit code 1 and the message on stderr when the reader throws.  Linked against host/file_io.cc only.
#include <cstdio>
#include <vector>

#include "../../gossamer_b200/host/file_io.hh"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    try {
        goss::InputFile in(argv[1]);
        std::vector<char> buf(70001);
        size_t want = 1;
        for (;;) {
            const size_t got = in.read(buf.data(), want);
            if (got) fwrite(buf.data(), 1, got, stdout);
            if (got < want) break;
            want = want * 3 + 1;
            if (want > buf.size()) want = buf.size();
        }
    } catch (const goss::Error& e) {
        fprintf(stderr, "%s", e.text.c_str());
        return 1;
    }
    return 0;
}
