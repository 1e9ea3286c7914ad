// sort_main.cpp -- the `sort` command line of the reference (lib/sort.h:28-120) in front of
// twkb_two_sort_mem (libtwkb.so): same options, range checks, messages and exit codes. -m is the memory budget per
// thread of the reference's external merge: inputs beyond it are sorted in spilled runs and merged k-way.
//
//   twkb_sort [sort] [options] -i <in.two> -o <out.two>
#include <getopt.h>

#include <cstdlib>
#include <iostream>
#include <string>
#include <thread>

#include "../../include/twkb.h"
#include "../../include/twkb_ld.hpp"

using twkb_host::timestamp;

static void sort_usage() {
    std::cerr << "About:  Sort TWO files\n\n"
                 "Usage:  twkb_sort sort [options] -i <in.two>\n\n"
                 "Options:\n"
                 "  -i FILE   input TWO file (required)\n"
                 "  -o FILE   output file (required)\n"
                 "  -m FLOAT  memory budget in GB per thread (default 0.5): larger inputs are sorted in runs spilled to temporary files and merged\n"
                 "  -c INT    compression level 1-20 (default: 1)\n"
                 "  -t INT    number of threads (default: maximum available)\n\n";
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "sort") { ++argv; --argc; }
    if (argc < 3) { sort_usage(); return 0; }
    static struct option long_options[] = {{"input", required_argument, 0, 'i'}, {"output", optional_argument, 0, 'o'},
                                           {"memory-usage", optional_argument, 0, 'm'}, {"compression-level", optional_argument, 0, 'c'},
                                           {"threads", optional_argument, 0, 't'}, {0, 0, 0, 0}};
    std::string in, out;
    float memory_limit = 0.5f;
    int c_level = 1, n_threads = (int)std::thread::hardware_concurrency();
    int c = 0, long_index = 0;
    while ((c = getopt_long(argc, argv, "i:o:m:c:t:?", long_options, &long_index)) != -1) {
        switch (c) {
            case 'i': in = optarg; break;
            case 'o': out = optarg; break;
            case 'm': memory_limit = (float)std::atof(optarg); break;
            case 'c': c_level = std::atoi(optarg); break;
            case 't': n_threads = std::atoi(optarg); break;
            default: std::fprintf(stderr, "%s: option `-%c' is invalid: ignored\n", argv[0], optopt); break;
        }
    }
    if (in.empty()) { std::cerr << timestamp("ERROR") << "No input value specified..." << std::endl; return 1; }
    if (memory_limit <= 0) { std::cerr << timestamp("ERROR") << "Cannot set memory limit <= 0..." << std::endl; return 1; }
    if (n_threads <= 0) { std::cerr << timestamp("ERROR") << "Cannot set number of threads <= 0..." << std::endl; return 1; }
    if (c_level <= 0) { std::cerr << timestamp("ERROR") << "Cannot set the compression level <= 0..." << std::endl; return 1; }
    if (out.empty() || out == "-") { std::cerr << timestamp("ERROR") << "Writing to stdout is not supported: give -o <output.two>" << std::endl; return 1; }
    std::cerr << timestamp("LOG") << "Calling sort..." << std::endl;
    char err[1024] = {0};
    uint64_t n = 0;
    // -m: GB per thread in the reference (two_sorter_settings::memory_limit, lib/sort.h); here the budget of the whole sort
    const uint64_t budget = (uint64_t)((double)memory_limit * 1e9 * (double)n_threads);
    const int rc = twkb_two_sort_mem(in.c_str(), out.c_str(), c_level, n_threads, budget, &n, err, sizeof(err));
    if (rc != TWKB_OK) { std::cerr << timestamp("ERROR") << err << std::endl; return 1; }
    std::cerr << timestamp("LOG") << "Sorted " << twkb_host::pretty(n) << " records..." << std::endl;
    std::cerr << timestamp("LOG") << "Finished!" << std::endl;
    return 0;
}
