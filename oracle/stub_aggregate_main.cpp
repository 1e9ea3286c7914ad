/* main() for oracle/_ref/tomahawk_aggregate: the reference's own two_reader::Aggregate (lib/two_reader.cpp:543-853, the engine
 * of the `aggregate` subcommand, lib/aggregate.h) behind a stub for lib/tomahawk.cpp. TEST INFRASTRUCTURE ONLY.
 *   tomahawk_aggregate aggregate -i <in.two> -f <r2|r|d|dprime|p|hets|alts> -r <mean|max|min|count|total|sd> -x X -y Y -c cutoff -t threads
 * stdout: "range bpx bpy n_original" then one line per contig "range min max", then X lines of Y values (%.17g) read as
 * data[i * X + j] -- the layout twk_agg_slave::Overload writes (lib/aggregation.h:208-219). */
#include "stub_common.h"
#include "utility.h"
#include "two_reader.h"
#include <getopt.h>
#include <cstdio>
int main(int argc, char** argv){
	if(argc < 2 || std::string(argv[1]) != "aggregate"){ std::cerr << "usage: tomahawk_aggregate aggregate -i in.two -f func -r reduce -x X -y Y -c cutoff -t threads" << std::endl; return 2; }
	tomahawk::twk_two_settings settings;
	settings.n_threads = 1;
	std::string f = "r2", r = "mean"; int32_t x = 1000, y = 1000, cutoff = 5; int c;
	while((c = getopt(argc - 1, argv + 1, "i:f:r:x:y:c:t:")) != -1){
		if(c == 'i') settings.in = optarg;
		else if(c == 'f') f = optarg;
		else if(c == 'r') r = optarg;
		else if(c == 'x') x = atoi(optarg);
		else if(c == 'y') y = atoi(optarg);
		else if(c == 'c') cutoff = atoi(optarg);
		else if(c == 't') settings.n_threads = atoi(optarg);
	}
	tomahawk::two_reader oreader;
	tomahawk::twk1_aggregate_t agg;
	if(!oreader.Aggregate(agg, settings, f, r, x, y, cutoff, false, false)) return 1;
	printf("%llu %u %u %u\n", (unsigned long long)agg.range, agg.bpx, agg.bpy, agg.n_original);
	printf("%u\n", (unsigned)agg.rid_offsets.size());
	for(size_t i = 0; i < agg.rid_offsets.size(); ++i) printf("%llu %u %u\n", (unsigned long long)agg.rid_offsets[i].range, agg.rid_offsets[i].min, agg.rid_offsets[i].max);
	for(int i = 0; i < x; ++i){
		for(int j = 0; j < y; ++j) printf(j ? " %.17g" : "%.17g", agg.data[i * x + j]);
		printf("\n");
	}
	return 0;
}
