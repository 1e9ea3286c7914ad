/* main() for oracle/_ref/tomahawk_scalc: the reference's own `scalc` subcommand
 * (lib/scalc.h:49) behind a stub for lib/tomahawk.cpp. TEST INFRASTRUCTURE ONLY. */
#include "stub_common.h"
#include "utility.h"
#include "scalc.h"
int main(int argc, char** argv){
	for(int i = 0; i < argc; ++i){ tomahawk::LITERAL_COMMAND_LINE += argv[i]; tomahawk::LITERAL_COMMAND_LINE += ' '; }
	if(argc < 2 || std::string(argv[1]) != "scalc"){ std::cerr << "usage: tomahawk_scalc scalc <args>" << std::endl; return 2; }
	return scalc(argc, argv);
}
