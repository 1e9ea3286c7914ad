// calc_main.cpp -- the `calc` command line of the reference (lib/calc.h:28-240) in front of the
// B200 engine: same options, same range checks, same messages and exit codes; everything is
// forwarded to twkb_host::twk_ld::Compute (include/twkb_ld.hpp), i.e. to libtwkb.so.
//
//   twkb_calc [calc] [options] -i <in.twk> -o <output.two>
//   twkb_calc scalc  [options] -i <in.twk> -I <contig:pos> -o <output.two>      (reference lib/scalc.h)
//
// Additions: -g/--devices LIST (CUDA ordinals, comma separated; one context per entry) and
// -K/--kernel auto|popc|umma|fp4, --host-unpack, --no-shards, --sorted.
#include <getopt.h>

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "../../include/twkb_ld.hpp"

using twkb_host::timestamp;

static void calc_usage() {
    std::cerr << "About:  Calculate linkage disequilibrium (B200 engine, libtwkb v" << twkb_version() << ")\n"
                 "        Force phased -p or unphased -u for faster calculations if\n"
                 "        all variant sites are guaranteed to have the given phasing.\n\n"
                 "Usage:  twkb_calc calc [options] -i <in.twk> -o <output.two>\n\n"
                 "Options:\n"
                 "  -i FILE   input Tomahawk (required)\n"
                 "  -o FILE   output file or file prefix (required)\n"
                 "  -t INT    number of host threads for unpacking (default: maximum available)\n"
                 "  -c INT    number of subproblems to split compute into (must be in (c!2 + c))\n"
                 "  -C INT    chosen part to compute (0 < -C < -c)\n"
                 "  -m        accepted for compatibility (low-memory mode of the CPU engine)\n"
                 "  -M        accepted for compatibility; triggers -m and -p like the reference\n"
                 "  -b        accepted for compatibility (input block size)\n"
                 "  -w INT    sliding window width in bases\n"
                 "  -I STRING filter interval <contig>:pos-pos (see manual)\n"
                 "  -p        force computations to use phased math\n"
                 "  -u        force computations to use unphased math\n"
                 "  -P FLOAT  Fisher's exact test / Chi-squared cutoff P-value (default: 1)\n"
                 "  -r FLOAT  Pearson's R-squared minimum cut-off value (default: 0.1)\n"
                 "  -k INT    compression level to use (default: 1, max = 22).\n"
                 "  -g LIST   CUDA devices, e.g. 0,1,2,3 (default: 0)\n"
                 "  -K NAME   count kernel: auto | popc | umma | fp4 (default: auto)\n"
                 "  --host-unpack  unpack the .twk genotypes on the host instead of decoding the runs on the device\n"
                 "  --no-shards    -w on several devices: deal tiles of the whole matrix instead of position shards\n"
                 "  --sorted       sort the records on the device and write a sorted, indexed .two (calc + sort in one run)\n"
              << std::endl;
}

static bool all_digits(const char* s) {
    if (!*s) return false;
    for (; *s; ++s)
        if (*s < '0' || *s > '9') return false;
    return true;
}

// "^[0-9]+$" -> atoi; "^[0-9]+[eE][0-9]+$" -> atof (lib/calc.h:198-211)
static bool parse_window(const char* s, int32_t* out) {
    if (all_digits(s)) { *out = std::atoi(s); return true; }
    const char* e = std::strpbrk(s, "eE");
    if (!e || e == s) return false;
    const std::string mant(s, e - s);
    if (!all_digits(mant.c_str()) || !all_digits(e + 1)) return false;
    *out = (int32_t)std::atof(s);
    return true;
}

// `scalc` (reference lib/scalc.h:28-194): one target site against its neighbourhood. Same option letters, range checks and
// messages; -w is the neighbourhood in bases (l_surrounding), -r keeps its twk_ld_settings default.
static void scalc_usage() {
    std::cerr << "About:  Calculate linkage disequilibrium for a single variant versus its neighbourhood (B200 engine).\n\n"
                 "Usage:  twkb_calc scalc [options] -i <in.twk> -I <SNP position> -o <output.two>\n\n"
                 "Options:\n"
                 "  -i FILE   input Tomahawk (required)\n"
                 "  -o FILE   output file or file prefix (required)\n"
                 "  -I STRING filter interval <contig>:pos-pos (see manual; required)\n"
                 "  -w INT    number of bases to include around the target snp (default: 500kbp)\n"
                 "  -t INT    number of host threads (default: maximum available)\n"
                 "  -m, -M, -b  accepted for compatibility\n"
                 "  -P FLOAT  Fisher's exact test / Chi-squared cutoff P-value (default: 1)\n"
                 "  -r FLOAT  Pearson's R-squared minimum cut-off value\n"
                 "  -k INT    compression level to use (default: 1, max = 22).\n"
                 "  -g INT    CUDA device (default: 0)\n"
              << std::endl;
}

static int scalc_main(int argc, char** argv) {
    if (argc < 3) {
        scalc_usage();
        return 1;
    }
    static struct option long_options[] = {{"input", required_argument, 0, 'i'},   {"output", required_argument, 0, 'o'},
                                           {"interval", required_argument, 0, 'I'}, {"window", optional_argument, 0, 'w'},
                                           {"threads", optional_argument, 0, 't'}, {"low-memory", optional_argument, 0, 'm'},
                                           {"block-size", optional_argument, 0, 'b'}, {"bitmaps", optional_argument, 0, 'M'},
                                           {"compression-level", optional_argument, 0, 'k'}, {"minP", optional_argument, 0, 'P'},
                                           {"minR2", optional_argument, 0, 'r'},   {"silent", no_argument, 0, 's'},
                                           {"devices", required_argument, 0, 'g'}, {0, 0, 0, 0}};
    twkb_host::twk_ld_settings settings;
    std::string literal;
    for (int i = 0; i < argc; ++i) literal += (i ? " " : "") + std::string(argv[i]);
    auto err = [](const char* m) {
        std::cerr << timestamp("ERROR") << m << std::endl;
        return 1;
    };
    int c, option_index = 0;
    while ((c = getopt_long(argc, argv, "i:o:t:P:a:A:r:I:smMb:k:w:g:?", long_options, &option_index)) != -1) {
        switch (c) {
            case 'i': settings.in = optarg; break;
            case 'o': settings.out = optarg; break;
            case 'I': settings.ival_strings.push_back(optarg); break;
            case 'm': settings.low_memory = true; break;
            case 'M': settings.force_phased = true; settings.low_memory = true; settings.bitmaps = true; break;
            case 't':
                settings.n_threads = std::atoi(optarg);
                if (settings.n_threads <= 0) return err("Cannot have a non-positive number of worker threads");
                break;
            case 'b':
                settings.bl_size = std::atoi(optarg);
                if (settings.bl_size <= 0) return err("Cannot have a non-positive number of entries in a block!");
                break;
            case 'r':
                settings.minR2 = std::atof(optarg);
                if (settings.minR2 < 0) return err("Cannot have a negative minimum R-squared value");
                if (settings.minR2 > 1) return err("Cannot have minimum R-squared value > 1");
                break;
            case 'P':
                settings.minP = std::atof(optarg);
                if (settings.minP < 0) return err("Cannot have a negative cutoff P-value");
                if (settings.minP > 1) return err("Cannot have a cutoff P-value > 1");
                break;
            case 'k': settings.c_level = std::atoi(optarg); break;
            case 'w':
                settings.l_surrounding = std::atoi(optarg);
                if (settings.l_surrounding < 1) return err("Cannot have neighbourhood (-w) <= 1");
                break;
            case 's': settings.silent = true; break;
            case 'a': case 'A': break;
            case 'g':
                if (!all_digits(optarg)) return err("Illegal device list");
                settings.devices.assign(1, std::atoi(optarg));
                break;
            default:
                std::cerr << timestamp("ERROR") << "Unrecognized option: " << (char)c << std::endl;
                return 1;
        }
    }
    if (settings.in.empty()) return err("No input value specified...");
    if (settings.out.empty()) return err("No output value specified...");
    if (!settings.silent) std::cerr << timestamp("LOG") << "Calling scalc..." << std::endl;
    settings.minR2 = 0;  // the reference's scalc overrides -r after parsing it (lib/scalc.h:188): every neighbour is reported
    twkb_host::twk_ld ld;
    ld.command_line = literal;
    return ld.ComputeSingle(settings) ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::strcmp(argv[1], "scalc") == 0) return scalc_main(argc - 1, argv + 1);
    if (argc >= 2 && std::strcmp(argv[1], "calc") == 0) { --argc; ++argv; }
    if (argc < 3) {
        calc_usage();
        return 1;
    }
    static struct option long_options[] = {{"input", required_argument, 0, 'i'},
                                           {"threads", optional_argument, 0, 't'},
                                           {"output", required_argument, 0, 'o'},
                                           {"interval", optional_argument, 0, 'I'},
                                           {"parts", optional_argument, 0, 'c'},
                                           {"partStart", optional_argument, 0, 'C'},
                                           {"low-memory", optional_argument, 0, 'm'},
                                           {"block-size", optional_argument, 0, 'b'},
                                           {"bitmaps", optional_argument, 0, 'M'},
                                           {"compression-level", optional_argument, 0, 'k'},
                                           {"cross-chr-only", no_argument, 0, 'X'},
                                           {"no-cross-chr", no_argument, 0, 'x'},
                                           {"minP", optional_argument, 0, 'P'},
                                           {"force-phased", no_argument, 0, 'p'},
                                           {"force-unphased", no_argument, 0, 'u'},
                                           {"samples", optional_argument, 0, 'S'},
                                           {"minR2", optional_argument, 0, 'r'},
                                           {"detailedProgress", no_argument, 0, 'd'},
                                           {"silent", no_argument, 0, 's'},
                                           {"windowBases", optional_argument, 0, 'w'},
                                           {"devices", required_argument, 0, 'g'},
                                           {"kernel", required_argument, 0, 'K'},
                                           {"host-unpack", no_argument, 0, 1001},
                                           {"no-shards", no_argument, 0, 1002},
                                           {"sorted", no_argument, 0, 1003},
                                           {0, 0, 0, 0}};
    twkb_host::twk_ld_settings settings;
    std::string literal;
    for (int i = 0; i < argc; ++i) literal += (i ? " " : "") + std::string(argv[i]);
    int c, option_index = 0;
    auto err = [](const char* m) {
        std::cerr << timestamp("ERROR") << m << std::endl;
        return 1;
    };
    while ((c = getopt_long(argc, argv, "i:o:t:puP:a:A:r:w:S:I:sdc:C:mMb:xXk:g:K:?", long_options, &option_index)) != -1) {
        switch (c) {
            case 'i': settings.in = optarg; break;
            case 'o': settings.out = optarg; break;
            case 'I': settings.ival_strings.push_back(optarg); break;
            case 'm': settings.low_memory = true; break;
            case 'p': settings.force_phased = true; settings.forced_unphased = false; break;
            case 'u': settings.forced_unphased = true; settings.force_phased = false; break;
            case 'M': settings.force_phased = true; settings.low_memory = true; settings.bitmaps = true; break;
            case 't':
                settings.n_threads = std::atoi(optarg);
                if (settings.n_threads <= 0) return err("Cannot have a non-positive number of worker threads");
                break;
            case 'b':
                settings.bl_size = std::atoi(optarg);
                if (settings.bl_size <= 0) return err("Cannot have a non-positive number of entries in a block!");
                break;
            case 'c':
                settings.n_chunks = std::atoi(optarg);
                if (settings.n_chunks <= 0) return err("Cannot have a negative or zero amount of partitions");
                break;
            case 'C':
                settings.c_chunk = std::atoi(optarg) - 1;
                if (settings.c_chunk < 0) return err("Cannot have a non-positive start partition");
                break;
            case 'r':
                settings.minR2 = std::atof(optarg);
                if (settings.minR2 < 0) return err("Cannot have a negative minimum R-squared value");
                if (settings.minR2 > 1) return err("Cannot have minimum R-squared value > 1");
                break;
            case 'P':
                settings.minP = std::atof(optarg);
                if (settings.minP < 0) return err("Cannot have a negative cutoff P-value");
                if (settings.minP > 1) return err("Cannot have a cutoff P-value > 1");
                break;
            case 'w':
                settings.window = true;
                if (!parse_window(optarg, &settings.l_window)) {
                    std::cerr << "not an integer" << std::endl;
                    return 1;
                }
                if (settings.l_window <= 0) return err("Cannot have a non-positive window size");
                break;
            case 'k': settings.c_level = std::atoi(optarg); break;
            case 's': settings.silent = true; break;
            case 'd': case 'x': case 'X': case 'S': case 'a': case 'A': break;  // parsed and ignored by the reference too
            case 'g': {
                settings.devices.clear();
                std::string list = optarg;
                size_t pos = 0;
                while (pos <= list.size()) {
                    const size_t comma = list.find(',', pos);
                    const std::string tok = list.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
                    if (!all_digits(tok.c_str())) return err("Illegal device list");
                    settings.devices.push_back(std::atoi(tok.c_str()));
                    if (comma == std::string::npos) break;
                    pos = comma + 1;
                }
                break;
            }
            case 1001: settings.host_unpack = true; break;
            case 1002: settings.position_shards = false; break;
            case 1003: settings.sorted_output = true; break;
            case 'K': {
                const std::string k = optarg;
                if (k == "auto") settings.kernel = TWKB_KERNEL_AUTO;
                else if (k == "popc") settings.kernel = TWKB_KERNEL_POPC;
                else if (k == "umma" || k == "i8") settings.kernel = TWKB_KERNEL_UMMA;
                else if (k == "fp4") settings.kernel = TWKB_KERNEL_UMMA_FP4;
                else return err("Unknown kernel (auto | popc | umma | fp4)");
                break;
            }
            default:
                std::cerr << timestamp("ERROR") << "Unrecognized option: " << (char)c << std::endl;
                return 1;
        }
    }
    if (settings.in.empty()) return err("No input value specified...");
    if (settings.out.empty()) return err("No output value specified...");
    if (!settings.silent) std::cerr << timestamp("LOG") << "Calling calc..." << std::endl;
    twkb_host::twk_ld ld;
    ld.command_line = literal;
    if (!ld.Compute(settings)) return 1;
    return 0;
}
