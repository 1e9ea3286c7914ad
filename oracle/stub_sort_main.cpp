/* main() for oracle/_ref/tomahawk_sort: the reference's own `sort` subcommand
 * (lib/sort.h -> two_reader::Sort, lib/two_reader.cpp:162-420). TEST INFRASTRUCTURE ONLY:
 * the checker of the product's .two sorter. */
#include "stub_common.h"
#include "utility.h"
#include "writer.h"
#include "sort.h"
int main(int argc, char** argv){
	if(argc < 2 || std::string(argv[1]) != "sort"){ std::cerr << "usage: tomahawk_sort sort <args>" << std::endl; return 2; }
	return sort(argc, argv);
}
