/* A C caller of the batched entry points, as INTEGRATION.md section 2 describes the move
 * from the reference's one-particle loop (examples/example-stepper.c:102-140): freeze the
 * stepper, trace a fan from one station asking for two columns, trace explicit rays into
 * records, walk particles for a few steps with their stepper state on the device.
 *
 *     batched_caller [n_azimuth n_elevation]
 *
 * Prints one line per stage. Without a CUDA device the first batched call fails through
 * the error handler (there is no CPU fallback): the handler prints the library's message
 * and the program exits with code 3. tests/test_batched_caller.py compiles this file with
 * -std=c99 -Wall -Werror against include/ and runs it. */
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "turtle.h"
#include "turtle_b200.h"

static void handle_error(enum turtle_return code, turtle_function_t * function,
    const char * message)
{
        (void)function;
        fprintf(stderr, "turtle error %d: %s\n", (int)code, message);
        exit(3);
}

int main(int argc, char * argv[])
{
        const size_t n_az = (argc > 1) ? (size_t)atol(argv[1]) : 64;
        const size_t n_el = (argc > 2) ? (size_t)atol(argv[2]) : 32;
        const size_t n = n_az * n_el;
        turtle_error_handler_set(&handle_error);

        /* a 101 x 101 UTM map with a ridge, over a flat sea */
        struct turtle_map * map;
        const struct turtle_map_info info = { 101, 101, { 486000., 488000. },
                { 5057000., 5059000. }, { 0., 2000. }, NULL };
        turtle_map_create(&map, &info, "UTM 31N");
        int ix, iy;
        for (ix = 0; ix < 101; ix++)
                for (iy = 0; iy < 101; iy++)
                        turtle_map_fill(map, ix, iy, 200. + 15. * (ix > 50 ? 100 - ix : ix));
        struct turtle_stepper * stepper;
        turtle_stepper_create(&stepper);
        turtle_stepper_range_set(stepper, 0.);
        turtle_stepper_add_flat(stepper, 0.);
        turtle_stepper_add_map(stepper, map, 0.);

        /* the station, as the reference's example places it */
        const double latitude = 45.675, longitude = 2.835;
        double position[3];
        turtle_stepper_position(stepper, latitude, longitude, 1., 0, position, NULL);
        printf("station %.3f %.3f %.3f\n", position[0], position[1], position[2]);
        fflush(stdout);

        /* 1. the plan: tiles and geometry go to the device once */
        struct turtle_plan * plan;
        turtle_stepper_freeze(stepper, 0, &plan);

        /* 2. a fan from the station; two columns come back */
        double * azimuth = malloc(n_az * sizeof(*azimuth));
        double * elevation = malloc(n_el * sizeof(*elevation));
        size_t i;
        for (i = 0; i < n_az; i++) azimuth[i] = 360. * (i + 0.5) / n_az;
        for (i = 0; i < n_el; i++) elevation[i] = 0.5 + 20. * i / n_el;
        struct turtle_fan fan = { latitude, longitude,
                { position[0], position[1], position[2] }, n_az, n_el, azimuth, elevation,
                32 };
        struct turtle_trace_rule rule = { -DBL_MAX, 3000., DBL_MAX, 100000, 0 };
        double * rock = malloc(n * sizeof(*rock));
        int32_t * status = malloc(n * sizeof(*status));
        struct turtle_trace_fields want;
        memset(&want, 0x0, sizeof(want)); /* NULL = column not wanted */
        want.length[0] = rock;
        want.status = status;
        turtle_stepper_trace_fan(plan, &fan, &rule, NULL, &want);
        double sum = 0.;
        for (i = 0; i < n; i++) sum += rock[i];
        printf("fan %zu rays, rock %.3f m in all\n", n, sum);

        /* 3. explicit rays into full records */
        double * origin = malloc(n * 3 * sizeof(*origin));
        double * direction = malloc(n * 3 * sizeof(*direction));
        for (i = 0; i < n; i++) {
                origin[3 * i] = position[0];
                origin[3 * i + 1] = position[1];
                origin[3 * i + 2] = position[2];
                turtle_ecef_from_horizontal(latitude, longitude, azimuth[i % n_az],
                    elevation[i / n_az], direction + 3 * i);
        }
        struct turtle_trace_result * results = malloc(n * sizeof(*results));
        turtle_stepper_trace_batch(plan, n, origin, direction, &rule, results);
        long steps = 0;
        for (i = 0; i < n; i++) steps += results[i].n_steps;
        printf("rays %zu, %ld steps\n", n, steps);

        /* 4. particles that turn every step: four steps per launch, state on the device */
        struct turtle_states * states;
        turtle_states_create(plan, n, &states);
        double * turns = malloc(4 * n * 3 * sizeof(*turns));
        int k;
        for (k = 0; k < 4; k++)
                for (i = 0; i < n; i++)
                        turtle_ecef_from_horizontal(latitude, longitude,
                            azimuth[(i + k) % n_az], elevation[i / n_az],
                            turns + 3 * (k * n + i));
        double * length = malloc(4 * n * sizeof(*length));
        int * index = malloc(4 * n * 2 * sizeof(*index));
        turtle_stepper_walk_batch(plan, states, n, 4, origin, turns, NULL, NULL, NULL, NULL,
            length, index);
        sum = 0.;
        for (i = 0; i < 4 * n; i++) sum += length[i];
        printf("walk %zu particles x 4 steps, %.3f m\n", n, sum);

        turtle_states_destroy(&states);
        turtle_plan_destroy(&plan);
        turtle_stepper_destroy(&stepper);
        turtle_map_destroy(&map);
        free(azimuth), free(elevation), free(rock), free(status), free(origin);
        free(direction), free(results), free(turns), free(length), free(index);
        return 0;
}
