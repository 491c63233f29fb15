// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A small driver of OUR OWN over the UNMODIFIED reference classes (linked from
// oracle/_ref/obj*, see oracle/Makefile.ref).  It exists because the stock
// `builder` binary (a) never exposes the raw BWT, (b) constructs RLCSABuilder
// with the default single thread (TextCollectionBuilder.cpp:55) and (c) never
// calls the dormant saveSamples() (FMIndex.cpp:125-147).  Nothing here is
// copied from the reference; it only calls its public interfaces:
//   CSA::RLCSABuilder(block, sample_rate, buffer, threads)  incbwt/rlcsa_builder.h:14
//   insertSequence / getBWT                                 incbwt/rlcsa_builder.h:18,30
//   FMIndex(bwt, n, samplerate, nTexts, maxLen, ...)         FMIndex.h / FMIndex.cpp:92
//   TextCollection::load / save / saveSamples               TextCollection.h:99-105
//
// Input "docs" file = the documents exactly as builder.cpp hands them to
// InsertText (already transformed), each terminated by one '\0' byte.
//
//   ref_driver bwt  <docs> <out.bwt>      [threads] [buffer_bytes]
//   ref_driver fmi  <docs> <out_prefix>   [threads] [buffer_bytes] [samplerate]
//   ref_driver sa   <index.fmi> <out_prefix>
//   ref_driver query <index.fmi> <queries.txt> <answers.txt>
//        one query per line: "L <symbol code> <i>" -> TextCollection::LF(c, i)  (FMIndex.h:84-90)
//                            "G <i>"               -> TextCollection::getL(i)   (FMIndex.h:99-102)
// Timings go to stdout as one line "seconds_total=... seconds_bwt=...".
#include "rlcsa_builder.h"
#include "TextCollection.h"
#include "FMIndex.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static double now()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

static std::vector<char> slurp(const char *path)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf(sz);
    if (sz && std::fread(buf.data(), 1, sz, f) != (size_t)sz) { std::perror("fread"); std::exit(2); }
    std::fclose(f);
    return buf;
}

struct Built { uchar *bwt; CSA::usint n; unsigned ntexts; ulong maxlen; double t_bwt; };

static Built build_bwt(const char *docs_path, CSA::usint threads, CSA::usint buffer)
{
    std::vector<char> docs = slurp(docs_path);
    Built b = {0, 0, 0, 0, 0.0};
    double t0 = now();
    // same parameters as TextCollectionBuilder.cpp:45-55 for TYPE_FMINDEX
    CSA::RLCSABuilder builder(CSA::RLCSA_BLOCK_SIZE.second, 0, buffer, threads);
    size_t pos = 0;
    while (pos < docs.size())
    {
        size_t len = std::strlen(docs.data() + pos);
        if (len == 0) { std::fprintf(stderr, "empty document at byte %zu\n", pos); std::exit(2); }
        builder.insertSequence(docs.data() + pos, len, false);
        b.ntexts++;
        if (len + 1 > b.maxlen) b.maxlen = len + 1;
        pos += len + 1;
    }
    CSA::usint length = 0;
    b.bwt = (uchar *)builder.getBWT(length);
    b.n = length;
    b.t_bwt = now() - t0;
    return b;
}

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: see header of oracle/ref_driver.cpp\n"); return 2; }
    std::string mode = argv[1];
    double t0 = now();
    if (mode == "bwt" || mode == "fmi")
    {
        CSA::usint threads = argc > 4 ? std::strtoul(argv[4], 0, 10) : 1;
        CSA::usint buffer = argc > 5 ? std::strtoul(argv[5], 0, 10) : (5lu * 1024 * 1024 * 1024) / 10;
        unsigned samplerate = argc > 6 ? std::strtoul(argv[6], 0, 10) : 124;
        Built b = build_bwt(argv[2], threads, buffer);
        if (!b.bwt) { std::fprintf(stderr, "getBWT failed\n"); return 1; }
        if (mode == "bwt")
        {
            FILE *f = std::fopen(argv[3], "wb");
            if (!f || std::fwrite(b.bwt, 1, b.n, f) != b.n) { std::perror(argv[3]); return 1; }
            std::fclose(f);
            delete[] b.bwt;
        }
        else
        {
            std::vector<std::string> names;
            TextCollection *tc = new FMIndex(b.bwt, (ulong)b.n, samplerate, b.ntexts, b.maxlen, 0, names, false, false, 0);
            tc->save(argv[3]);
            delete tc;
        }
        std::printf("seconds_total=%.3f seconds_bwt=%.3f n=%lu texts=%u threads=%lu\n",
                    now() - t0, b.t_bwt, (unsigned long)b.n, b.ntexts, (unsigned long)threads);
        return 0;
    }
    if (mode == "sa")
    {
        TextCollection *tc = TextCollection::load(argv[2]);
        tc->saveSamples(argv[3]);
        delete tc;
        std::printf("seconds_total=%.3f\n", now() - t0);
        return 0;
    }
    if (mode == "query")
    {
        if (argc < 5) { std::fprintf(stderr, "query needs <index.fmi> <queries> <answers>\n"); return 2; }
        TextCollection *tc = TextCollection::load(argv[2]);
        FILE *q = std::fopen(argv[3], "r"), *a = std::fopen(argv[4], "w");
        if (!q || !a) { std::perror("query files"); return 2; }
        char op;
        while (std::fscanf(q, " %c", &op) == 1)
        {
            unsigned long c = 0, i = 0;
            if (op == 'L')
            {
                if (std::fscanf(q, "%lu %lu", &c, &i) != 2) return 2;
                std::fprintf(a, "%lu\n", (unsigned long)tc->LF((uchar)c, (ulong)i));
            }
            else if (op == 'G')
            {
                if (std::fscanf(q, "%lu", &i) != 1) return 2;
                std::fprintf(a, "%u\n", (unsigned)tc->getL((ulong)i));
            }
            else
                return 2;
        }
        std::fclose(q);
        std::fclose(a);
        delete tc;
        return 0;
    }
    std::fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
}
