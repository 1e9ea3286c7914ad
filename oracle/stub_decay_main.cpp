/* main() for oracle/_ref/tomahawk_decay: the reference's own two_reader::Decay (lib/two_reader.cpp:424-475; the
 * `decay` subcommand of lib/decay.h currently routes to PositionalDecay and leaves this call commented out, :107)
 * behind a stub for lib/tomahawk.cpp. TEST INFRASTRUCTURE ONLY.
 *   tomahawk_decay decay -i <in.two> -w <window bp> -b <bins>   ->  "From To Mean Frequency" table on stdout */
#include "stub_common.h"
#include "utility.h"
#include "two_reader.h"
#include <getopt.h>
int main(int argc, char** argv){
	if(argc < 2 || std::string(argv[1]) != "decay"){ std::cerr << "usage: tomahawk_decay decay -i in.two -w bp -b bins" << std::endl; return 2; }
	tomahawk::twk_two_settings settings;
	int64_t window = 10000000; int32_t bins = 1000; int c;
	while((c = getopt(argc - 1, argv + 1, "i:w:b:")) != -1){
		if(c == 'i') settings.in = optarg;
		else if(c == 'w') window = atoll(optarg);
		else if(c == 'b') bins = atoi(optarg);
	}
	tomahawk::two_reader oreader;
	return oreader.Decay(settings, window, bins) ? 0 : 1;
}
