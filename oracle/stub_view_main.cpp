/* main() for oracle/_ref/tomahawk_view: the reference's own `view` subcommand
 * (lib/view.h) used as a .two -> text dumper. TEST INFRASTRUCTURE ONLY. */
#include "stub_common.h"
#include "utility.h"
#include "writer.h"
#include "view.h"
int main(int argc, char** argv){
	if(argc < 2 || std::string(argv[1]) != "view"){ std::cerr << "usage: tomahawk_view view <args>" << std::endl; return 2; }
	return view(argc, argv);
}
