/*
 * A caller written ONLY against the reference's public interface (turtle.h): the ray
 * loop of examples/example-stepper.c:102-140 on an in-memory map + flat layer. It is
 * compiled by tests/test_dropin.py twice, against the reference library and against
 * libturtle_b200.so, without any source change: the two runs must print the same
 * bits. argv[1] = number of rays.
 */
#include <stdio.h>
#include <stdlib.h>
#include "turtle.h"

static void on_error(enum turtle_return code, turtle_function_t * function,
    const char * message)
{
        fprintf(stderr, "turtle error: %s\n", message);
        exit(EXIT_FAILURE);
}

int main(int argc, char * argv[])
{
        const int n_rays = (argc > 1) ? atoi(argv[1]) : 16;
        turtle_error_handler_set(&on_error);

        /* a 101 x 101 UTM map with an analytic relief */
        struct turtle_map * map = NULL;
        struct turtle_map_info info = { 101, 101, { 486000., 496000. },
                { 5057000., 5067000. }, { 0., 3000. }, NULL };
        turtle_map_create(&map, &info, "UTM 31N");
        for (int iy = 0; iy < 101; iy++)
                for (int ix = 0; ix < 101; ix++)
                        turtle_map_fill(map, ix, iy,
                            1000. + 8. * ix + 3. * iy + 0.05 * (ix - 50) * (iy - 40));

        struct turtle_stepper * stepper = NULL;
        turtle_stepper_create(&stepper);
        turtle_stepper_range_set(stepper, 10.);
        turtle_stepper_add_flat(stepper, -100.);
        turtle_stepper_add_layer(stepper);
        turtle_stepper_add_map(stepper, map, 0.);

        const struct turtle_projection * projection = turtle_map_projection(map);
        double latitude, longitude;
        turtle_projection_unproject(projection, 491000., 5062000., &latitude, &longitude);

        for (int i = 0; i < n_rays; i++) {
                double position[3], direction[3], altitude, step;
                int index[2];
                turtle_stepper_reset(stepper);
                turtle_stepper_position(stepper, latitude, longitude, -5., 1, position, index);
                turtle_ecef_from_horizontal(latitude, longitude, 360. * i / n_rays,
                    2. + 0.5 * i, direction);
                turtle_stepper_step(stepper, position, NULL, NULL, NULL, &altitude, NULL, NULL,
                    index);
                double rock_length = 0.;
                int n_steps = 0;
                while ((altitude < 2500.) && (index[0] >= 0) && (n_steps < 100000)) {
                        const int initial_layer = index[0];
                        turtle_stepper_step(stepper, position, direction, NULL, NULL,
                            &altitude, NULL, &step, index);
                        if (initial_layer == 0) rock_length += step;
                        n_steps++;
                }
                printf("%d %d %d %a %a %a %a %a\n", i, n_steps, index[0], rock_length, altitude,
                    position[0], position[1], position[2]);
        }
        const char * name;
        turtle_map_meta(map, &info, &name);
        printf("map %d %d %a %a %s\n", info.nx, info.ny, info.x[1], info.z[1], name);
        turtle_stepper_destroy(&stepper);
        turtle_map_destroy(&map);
        return 0;
}
