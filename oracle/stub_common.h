/* Replaces the reference's lib/tomahawk.cpp (which needs htslib/hts.h, absent
 * in this image) with the handful of free functions/globals the calc and view
 * subcommands reference. TEST INFRASTRUCTURE ONLY. */
#include <iostream>
#include <string>
#include "tomahawk.h"
namespace tomahawk {
std::string LITERAL_COMMAND_LINE;
std::string INTERPRETED_COMMAND;
std::string LibrariesString(){ return std::string("Libraries: tomahawk-oracle"); }
void ProgramMessage(const bool separator){ if(separator) std::cerr << "----------" << std::endl; }
void ProgramHelp(void){ std::cerr << "oracle build: calc / view only" << std::endl; }
void ProgramHelpDetailed(void){ ProgramHelp(); }
}
