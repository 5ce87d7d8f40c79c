// dropin_tutorial.cpp -- a C++ user program over nlopt_mini.hpp (a wrapper in the style of the reference's nlopt.hpp)
// linked with -lnlopt against this repository's libnlopt.so.1: the flow of the reference's test/t_tutorial.cxx:37-82
// (algorithm from argv, parameters through set_param, functor data owned through the munge callbacks, a copy of the
// object that must own duplicates of the functor data, an exception thrown inside a callback).
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "nlopt_mini.hpp"

int main(int argc, char **argv)
{
    const nlopt_algorithm alg = argc > 1 ? (nlopt_algorithm) std::atoi(argv[1]) : NLOPT_LD_MMA;
    const double exactmin = 0.544331053951817355154952;   // sqrt(8/27)
    int count = 0;
    nlopt_mini::opt opt(alg, 2);
    opt.set_lower_bounds({-HUGE_VAL, 1e-6});
    opt.set_min_objective([&count](const std::vector<double> &x, std::vector<double> &g) {
        ++count;
        if (!g.empty()) { g[0] = 0.0; g[1] = 0.5 / std::sqrt(x[1]); }
        return std::sqrt(x[1]);
    });
    auto cons = [](double a, double b) {
        return [a, b](const std::vector<double> &x, std::vector<double> &g) {
            const double t = a * x[0] + b;
            if (!g.empty()) { g[0] = 3 * a * t * t; g[1] = -1.0; }
            return t * t * t - x[1];
        };
    };
    opt.add_inequality_constraint(cons(2, 0), 1e-8);
    opt.add_inequality_constraint(cons(-1, 1), 1e-8);
    opt.set_xtol_rel(1e-4);
    opt.set_param("inner_maxeval", 123);
    if (opt.get_param("inner_maxeval", 1234) != 123 || opt.get_param("not a param", 1234) != 1234) return 3;
    opt.set_param("rho_init", 0.5);

    try {
        std::vector<double> x = {1.234, 5.678};
        double minf = 0.0;
        opt.optimize(x, minf);
        std::cout << opt.get_algorithm_name() << " found minimum at f(" << x[0] << "," << x[1] << ") = " << std::setprecision(10)
                  << minf << " = exactmin + " << minf - exactmin << " after " << count << " evaluations" << std::endl;
        if (!(std::fabs(minf - exactmin) < 1e-3)) return 1;

        // a copy owns duplicates of the functor records (dup through the munge callback) and gives the same answer
        double minf2 = 0.0;
        std::vector<double> x2 = {1.234, 5.678};
        {
            nlopt_mini::opt twin(opt);
            twin.optimize(x2, minf2);
        }                                                  // twin destroyed: its duplicates freed, ours untouched
        if (minf2 != minf || x2 != x) { std::cerr << "copy gave a different result" << std::endl; return 4; }

        // an exception inside a callback stops the run and is re-thrown
        nlopt_mini::opt bad(alg, 2);
        bad.set_lower_bounds({-HUGE_VAL, 1e-6});
        int calls = 0;
        bad.set_min_objective([&calls](const std::vector<double> &x, std::vector<double> &g) -> double {
            if (++calls == 3) throw std::runtime_error("stop here");
            if (!g.empty()) { g[0] = 0.0; g[1] = 0.5 / std::sqrt(x[1]); }
            return std::sqrt(x[1]);
        });
        bad.add_inequality_constraint(cons(2, 0), 1e-8);
        bool thrown = false;
        try {
            std::vector<double> x3 = {1.234, 5.678};
            double m3;
            bad.optimize(x3, m3);
        } catch (const std::runtime_error &e) {
            thrown = std::string(e.what()) == "stop here";
        }
        if (!thrown) { std::cerr << "exception was not propagated" << std::endl; return 5; }
    } catch (std::exception &e) {
        std::cerr << "nlopt failed: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
