// prm_check.cc -- CPU-only tool used by the tests: parses a .prm file with the
// reference's declared entries and prints the resolved values and the
// expressions K reg(h), Eps reg(h), Pressure(time) evaluated at given points.
#include <cstdlib>
#include <iostream>

#include "fracture_problem.h"

int
main (int argc, char *argv[])
{
  using namespace cracks;
  if (argc < 2)
    return 2;
  try
    {
      ParameterHandler prm;
      FracturePhaseFieldProblem::declare_parameters (prm);
      prm.parse_input (argv[1]);
      const double h = argc > 2 ? std::atof (argv[2]) : 1.0, t = argc > 3 ? std::atof (argv[3]) : 1.0;
      prm.enter_subsection ("Problem dependent parameters");
      FunctionParser f;
      std::cout.precision (17);
      f.initialize ("h", prm.get ("K reg"));
      std::cout << "K=" << f.value (h) << "\n";
      f.initialize ("h", prm.get ("Eps reg"));
      std::cout << "Eps=" << f.value (h) << "\n";
      f.initialize ("time", prm.get ("Pressure"));
      std::cout << "Pressure=" << f.value (t) << "\n";
      prm.leave_subsection ();
      std::cout << prm.print_parameters ();
    }
  catch (std::exception &e)
    {
      std::cerr << e.what () << std::endl;
      return 1;
    }
  return 0;
}
