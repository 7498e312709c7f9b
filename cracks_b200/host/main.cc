// main.cc -- command line of the B200 build: `cracks_b200 <parameter_file>`
// like `./cracks <parameter_file>` (cracks.cc:4585-4686).  Without arguments it
// writes default.prm and prints the usage line, as the reference does after
// its unit tests.  Exit code 1 on any exception (cracks.cc:4662-4683).
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sys/stat.h>

#include "fracture_problem.h"

int
main (int argc, char *argv[])
{
  using namespace cracks;
  try
    {
      ParameterHandler prm;
      FracturePhaseFieldProblem::declare_parameters (prm);
      const char *file = nullptr;
      int device = 0, gmres_max_it = 200;
      std::string source_dir = ".";
      bool adaptive = false, write_output = true;
      for (int i = 1; i < argc; ++i)
        {
          if (!std::strcmp (argv[i], "--device") && i + 1 < argc)
            device = std::atoi (argv[++i]);
          else if (!std::strcmp (argv[i], "--gmres-max-it") && i + 1 < argc)
            gmres_max_it = std::atoi (argv[++i]);
          else if (!std::strcmp (argv[i], "--no-output"))
            write_output = false; // skip the .vtu files of output_results() (benchmark-sized runs)
          else if (!std::strcmp (argv[i], "--adaptive"))
            adaptive = true; // follow refine_mesh() on the host forest
          else if (!std::strcmp (argv[i], "--source-dir") && i + 1 < argc)
            source_dir = argv[++i]; // where test.pgm lives (multiple het)
          else
            file = argv[i];
        }
      if (!file)
        {
          std::ofstream out ("default.prm");
          out << prm.print_parameters ();
          std::cout << "usage: ./cracks_b200 <parameter_file> [--device N] [--gmres-max-it K] [--source-dir DIR] [--adaptive] [--no-output]" << std::endl
                    << " (created default.prm)" << std::endl;
          return 0;
        }
      prm.parse_input (file);
      prm.enter_subsection ("Global parameters");
      const std::string output_folder = prm.get ("Output directory");
      const int dim = (int) prm.get_integer ("Dimension");
      prm.leave_subsection ();
      ::mkdir (output_folder.c_str (), 0755);
      {
        std::ofstream out ((output_folder + "/parameters.prm").c_str ());
        out << prm.print_parameters ();
      }
      std::cout << "Problem dimension: " << dim << std::endl;
      FracturePhaseFieldProblem problem (prm, dim, std::cout);
      problem.device = device;
      problem.gmres_max_iterations = gmres_max_it;
      problem.source_dir = source_dir;
      problem.adaptive_forest = adaptive;
      problem.write_output = write_output;
      problem.run ();
    }
  catch (std::exception &exc)
    {
      std::cerr << std::endl
                << "----------------------------------------------------" << std::endl
                << "Exception on processing: " << std::endl
                << exc.what () << std::endl
                << "Aborting!" << std::endl
                << "----------------------------------------------------" << std::endl;
      return 1;
    }
  return 0;
}
