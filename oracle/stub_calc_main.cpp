/* main() for oracle/_ref/tomahawk_calc: the reference's own `calc` subcommand
 * (lib/calc.h:56) behind a stub for lib/tomahawk.cpp. TEST INFRASTRUCTURE ONLY. */
#include "stub_common.h"
#include "utility.h"
#include "calc.h"
int main(int argc, char** argv){
	for(int i = 0; i < argc; ++i){ tomahawk::LITERAL_COMMAND_LINE += argv[i]; tomahawk::LITERAL_COMMAND_LINE += ' '; }
	if(argc < 2 || std::string(argv[1]) != "calc"){ std::cerr << "usage: tomahawk_calc calc <args>" << std::endl; return 2; }
	return calc(argc, argv);
}
